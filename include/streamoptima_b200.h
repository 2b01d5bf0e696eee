/*
 * streamoptima_b200.h -- C ABI of the B200-native (sm_100a) per-block encode hot path of StreamOptima.
 *
 * The reference (Suyashagarw/StreamOptima) is pure Python and has no FFI layer; its boundary for this path is the
 * Python class surface `Y_Video_codec` (Encoder.py:24, :1790, :1544).  This header declares what the two per-frame
 * flows of that class call into once they are replaced:
 *
 *     Encoder.py:1582  complete_intra_flow   ->  so_encode_intra
 *     Encoder.py:1644  complete_inter_flow   ->  so_encode_inter
 *     Encoder.py:1790  encode() frame loop   ->  so_encode_sequence   (host buffers in / out, H2D + D2H inside)
 *     Encoder.py:1864  reference-list FIFO   ->  so_ref_reset / so_ref_push
 *     Encoder.py:1419  differential_encoder_frame  -> so_format_mv_frame      (host, text)
 *     Encoder.py:1522  entropy_encoder_frame       -> so_format_residual_frame (host, text)
 *
 * Conventions: plain C symbols; every call returns 0 on success or a negative SO_E_* code, with a message available
 * from so_last_error(); no exceptions or Python objects cross the ABI; frame buffers are caller-owned; a context is
 * bound to one CUDA device and is not thread-safe (one context per GPU worker process); device work is enqueued on
 * the caller's cudaStream_t (passed as void*) and is asynchronous unless stated otherwise.
 *
 * There is no CPU fallback: every entry point that computes fails with SO_E_CUDA when no sm_100 device is present.
 */
#ifndef STREAMOPTIMA_B200_H
#define STREAMOPTIMA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SO_ABI_VERSION 2

/* error codes */
#define SO_OK          0
#define SO_E_INVALID  -1   /* bad argument / unsupported parameter combination      */
#define SO_E_CUDA     -2   /* CUDA runtime / driver error (message has the detail)   */
#define SO_E_NOMEM    -3
#define SO_E_STATE    -4   /* call sequence error (e.g. inter frame with empty ring)  */
#define SO_E_RC       -5   /* no QP in the rate table satisfies the row budget (the reference crashes here, Encoder.py:1576) */

/* so_params.flags */
#define SO_FLAG_FME     1u   /* FMEEnable : half-pel search (Encoder.py:388, :697)     */
#define SO_FLAG_FAST_ME 2u   /* fast_me   : 3x3 around the predictor (Encoder.py:719)  */
#define SO_FLAG_VBS     4u   /* VBSEnable : 4-way split with RD decision (Encoder.py:512-573) */
#define SO_FLAG_SEA     8u   /* extension, results unchanged: the exhaustive search of find_best_match (Encoder.py:678-717)
                              * skips candidates whose SAD lower bound (8x8 quadrant sums) exceeds the exact SAD of a predictor
                              * candidate -- same minimum key, ties included.  Applies to 16x16 blocks at r = 16 without VBS;
                              * ignored elsewhere.  Off by default. */
#define SO_FLAG_SEA_AUTO 16u /* with SO_FLAG_SEA: the library watches how many exact SADs the pruned launches that have FINISHED took
                              * per (block, reference, phase plane) (read from mapped host memory, no synchronisation) and runs the
                              * plain kernel for the next 60 P frames whenever their running average says pruning does not pay on the
                              * content, then probes again.  The host enqueues a resident sequence far ahead of the GPU, so the
                              * feedback acts across calls (GOP after GOP, sequence after sequence), not inside one short call.
                              * Results are identical either way. */

#define SO_MAX_REF 8
#define SO_ALL_UNITS (-1)    /* `unit` argument of the per-frame calls: every unit of the context in lock step */

/* Encoder parameters; mirrors the Y_Video_codec constructor (Encoder.py:24). */
typedef struct so_params {
    int32_t width;          /* w_pixels, multiple of block_size (Encoder.py:1382)              */
    int32_t height;         /* h_pixels, multiple of block_size                                */
    int32_t block_size;     /* i : 4, 8 or 16                                                  */
    int32_t search_range;   /* r : integer-pel range 0..63; half-pel search covers +-2r (Encoder.py:1649) */
    int32_t qp;             /* Qp (const_init_Qp); 0 .. log2(block_size)+7                      */
    int32_t intra_dur;      /* I_Period                                                        */
    int32_t n_ref_frames;   /* nRefFrames, 1 .. SO_MAX_REF                                     */
    uint32_t flags;         /* SO_FLAG_*                                                       */
    int32_t rc_flag;        /* RCFlag: 0 off, 1 row-level table RC, 2 = 1 + scene-cut re-encode */
    int32_t parallel_mode;  /* ParallelMode 0, 1 or 2 (3 is broken in the reference)           */
    double  lam;            /* lambda of the RD cost (Encoder.py:1158); used only with VBS     */
    int64_t intra_thresh;   /* scene-cut threshold on quantized_sized (Encoder.py:1852)        */
    int32_t max_batch;      /* independent units (streams / closed GOPs) encoded per launch    */
    int32_t reserved;
} so_params;

/* Per-frame statistics written by the encode calls (device or host memory, see each call). */
typedef struct so_frame_stats {
    uint64_t sse;           /* sum of squared error recon vs input -> PSNR (Encoder.py:1869)   */
    uint64_t mae_num;       /* numerator of "MAE per Frame" (Encoder.py:583); see mae_den      */
    uint32_t mae_den;       /* average_mae = (mae_num / mae_den) / n_blocks                    */
    uint32_t mae_inf;       /* !=0: some block had no valid candidate -> average_mae = inf (quirk Q3) */
    uint32_t qsize;         /* quantized_sized: sum of RLE symbol counts (Encoder.py:1614,1683) */
    uint32_t frame_type;    /* 0 intra, 1 inter (after the scene-cut decision)                 */
} so_frame_stats;

/* Output planes of one frame.  All pointers are DEVICE pointers for so_encode_intra/inter.
 *   split  u8  [n_blocks]            0 whole block, 1 four sub-blocks (Z order)
 *   mv     i16 [n_blocks][4][3]      P: (dx,dy,ref) in slot 0 or the four sub-block vectors; I: offset in [k][0]
 *   levels i16 [height][width]       quantised coefficients at the pixel position of their (sub-)block
 *   recon  u8  [height][width]       reconstructed frame
 *   row_sizes u32 [height/block_size]  RLE symbols per block row (bits_spent_per_row, Encoder.py:1627)
 */
typedef struct so_frame_out {
    uint8_t*  split;
    int16_t*  mv;
    int16_t*  levels;
    uint8_t*  recon;
    uint32_t* row_sizes;
    so_frame_stats* stats;
} so_frame_out;

/* Packed run-level symbols of a sequence as a HOST output of the sequence encodes (so_set_symbol_output): the symbol lists
 * entropy_encoder_block (Encoder.py:1086-1131) produces, int16, frame after frame.  A frame is the concatenation of its
 * blocks' lists in raster order (the four lists of a split block in Z order); lists are self-delimiting -- a list ends
 * with the symbol 0 or when its runs have covered the block (decoder.py:548-586).  They replace the 2 B/px raw levels on
 * the device-to-host path: the residual text is formatted from them (so_write_bitstream_files_symbols) and the levels
 * can be rebuilt on the host when wanted (so_symbols_to_levels). */
typedef struct so_symbol_out {
    int16_t*  symbols;      /* host buffer (pinned for asynchronous copies), `capacity` symbols                      */
    uint64_t  capacity;
    uint64_t* pos;          /* out [n_units][n_frames]: index of the frame's first symbol in `symbols`                */
    uint32_t* count;        /* out [n_units][n_frames]: symbols of the frame                                         */
    uint64_t  needed;       /* out: symbols of the whole sequence (> capacity: SO_E_NOMEM, see so_fetch_symbols)      */
} so_symbol_out;

typedef struct so_ctx so_ctx;

int         so_abi_version(void);
const char* so_last_error(const so_ctx* ctx);          /* ctx may be NULL: error of the last failed create */
int         so_device_count(void);                     /* number of visible sm_100 devices, or SO_E_CUDA  */

int  so_ctx_create(so_ctx** out, const so_params* p, int device);
void so_ctx_destroy(so_ctx* ctx);

/* Change Qp / const_init_Qp for subsequent encodes (set_Qp, Encoder.py:948). */
int so_set_qp(so_ctx* ctx, int qp);

/* Row QPs for rate control (RCFlag>0): data-independent (Encoder.py:1599-1609), computed by the host wrapper and
 * shared by every frame.  n must equal height/block_size. */
int so_set_row_qps(so_ctx* ctx, const int32_t* qp_rows, int n);

/* ROI extension -- NOT part of the reference (its README advertises ROI, its code has none): per-block QPs
 * for the final quantisation of the next so_encode_sequence / so_seq_run / so_decode_sequence calls, i32
 * [n_frames][n_blocks], shared by all units; NULL clears.  The reference text format carries a QP only at block-row
 * starts, so a per-block map is side information the unchanged decoder.py cannot read. */
int so_set_block_qps(so_ctx* ctx, const int32_t* qp_blocks, int n_frames);

/* Reference list (Encoder.py:1798, :1864-1867).  A context holds max_batch independent chains ("units": streams or closed
 * GOPs).  unit = 0 .. max_batch-1 addresses ONE chain: the frame buffers of the call are single frames.  unit =
 * SO_ALL_UNITS addresses every chain in lock step: frame buffers are dense [max_batch][...] arrays.  Mixing is allowed
 * in one direction -- after lock-step calls every unit continues from the shared state; going back to SO_ALL_UNITS
 * needs identical chains or so_ref_reset(ctx, SO_ALL_UNITS, ..), else SO_E_STATE. */
int so_ref_reset(so_ctx* ctx, int unit, void* stream);                 /* list := [constant-128 float frame] */
int so_ref_push(so_ctx* ctx, int unit, const uint8_t* recon_dev, void* stream);   /* FIFO append of uint8 frame(s) [height][width] */

/* One frame: cur_dev is u8 [height][width] on the device (unit >= 0) or [max_batch][height][width] (SO_ALL_UNITS); the
 * planes of `out` follow the same rule.  Asynchronous on `stream`.  The type of the frame is the caller's decision
 * (Encoder.py:1839).  so_encode_inter does NOT push the reconstruction into the list (Encoder.py:1864-1867 is the
 * caller's so_ref_push), which is what makes teacher-forced parity tests possible (tests/test_gpu_seam.py). */
int so_encode_intra(so_ctx* ctx, int unit, const uint8_t* cur_dev, const so_frame_out* out, void* stream);
int so_encode_inter(so_ctx* ctx, int unit, const uint8_t* cur_dev, const so_frame_out* out, void* stream);

/* Whole-sequence encode of `n_units` independent sequences of `n_frames` frames each, HOST buffers in and out
 * (pinned or pageable).  Restates the frame loop of Encoder.py:1829-1871 including frame typing, the reference
 * FIFO, ParallelMode 1/2 semantics and the RCFlag=2 scene-cut re-encode.  Synchronous.
 *   frames     u8  [n_units][n_frames][height][width]
 *   split      u8  [n_units][n_frames][n_blocks]
 *   mv         i16 [n_units][n_frames][n_blocks][4][3]
 *   levels     i16 [n_units][n_frames][height][width]      (may be NULL: not copied back)
 *   recon      u8  [n_units][n_frames][height][width]      (may be NULL)
 *   row_sizes  u32 [n_units][n_frames][height/block_size]  (may be NULL)
 *   stats          [n_units][n_frames]
 */
int so_encode_sequence(so_ctx* ctx, const uint8_t* frames, int n_units, int n_frames,
                       uint8_t* split, int16_t* mv, int16_t* levels, uint8_t* recon,
                       uint32_t* row_sizes, so_frame_stats* stats);

/* Arm (sym != NULL) or disarm (NULL) symbol output for the following so_encode_sequence / so_encode_yuv420_file calls of
 * this context.  While armed, every chunk of frames is run through count -> device prefix scan -> emit on the device as
 * soon as it is encoded and its symbols are copied into sym->symbols with their exact sizes, overlapped with the encode
 * of the next chunk (frames are stored in the order they finish: use pos / count).  `levels` may then be NULL in those
 * calls.  The struct must stay alive while armed.  If the buffer is too small the call still completes every other
 * output, sets sym->needed and returns SO_E_NOMEM; the symbols stay resident on the device:
 * so_fetch_symbols(ctx, sym) with a larger buffer copies them (in unit, frame order) without encoding again. */
int so_set_symbol_output(so_ctx* ctx, so_symbol_out* sym);
int so_fetch_symbols(so_ctx* ctx, so_symbol_out* sym);

/* Frame ingest fused with the encode of ONE sequence (read_yuv Encoder.py:110-126, pad_hw :140-155): frames
 * first_frame .. first_frame + n_frames - 1 of a planar YUV 4:2:0 file whose luma is src_width x src_height (<= the coded
 * size of the context; the missing rows / columns are padded with 128 like pad_hw does).  Only the luma planes are read,
 * in chunks of 8 frames into rotating pinned buffers: disk reads, H2D copies, the encode and the D2H copies of different
 * chunks overlap.  Outputs as so_encode_sequence with n_units = 1.  Synchronous. */
int so_encode_yuv420_file(so_ctx* ctx, const char* path, int src_width, int src_height, int first_frame, int n_frames,
                          uint8_t* split, int16_t* mv, int16_t* levels, uint8_t* recon, uint32_t* row_sizes,
                          so_frame_stats* stats);

/* The three stages of so_encode_sequence, separately callable (bench.py times so_seq_run alone with the inputs
 * already resident in HBM, and the whole so_encode_sequence for the end-to-end figure):
 *   so_seq_upload   H2D copy of the frames, asynchronous on the context stream
 *   so_seq_run      the frame loop; outputs stay on the device; asynchronous unless rc_flag == 2
 *   so_seq_download D2H copy of the requested outputs + stream synchronise
 *   so_seq_sync     stream synchronise only */
int so_seq_upload(so_ctx* ctx, const uint8_t* frames, int n_units, int n_frames);
int so_seq_run(so_ctx* ctx);
int so_seq_download(so_ctx* ctx, uint8_t* split, int16_t* mv, int16_t* levels, uint8_t* recon,
                    uint32_t* row_sizes, so_frame_stats* stats);
int so_seq_sync(so_ctx* ctx);

/* Run-level symbols (entropy_encoder_block, Encoder.py:1086-1131) of the sequence resident after so_seq_run /
 * so_encode_sequence, generated on the device: per-(sub-)block counts, a device-side exclusive prefix scan, then every
 * symbol is written at its final position of a packed per-frame int16 stream.
 *   so_seq_symbols           runs the three kernels (asynchronous)
 *   so_seq_download_symbols  offsets u32 [units*frames][4*n_blocks + 1] (entry [b*4+k] = start of sub-block k of block b
 *                            inside its frame, last entry = symbols in the frame); symbols i16 packed, frame f starting
 *                            at sym_base[f] (u64 [units*frames + 1]); *needed = total symbols.  SO_E_NOMEM when
 *                            sym_capacity is too small (call again with `*needed`).
 *   so_format_residual_frame_symbols  the residual text of one frame from its symbols (same bytes as so_format_residual_frame) */
int so_seq_symbols(so_ctx* ctx);
int so_seq_download_symbols(so_ctx* ctx, uint32_t* offsets, int16_t* symbols, uint64_t sym_capacity, uint64_t* sym_base, uint64_t* needed);
int64_t so_format_residual_frame_symbols(const uint8_t* split, const uint32_t* offsets, const int16_t* symbols, int n_blocks,
                                         char* dst, int64_t cap);

/* Host side of the packed symbol streams (the so_symbol_out layout); no device needed:
 *   so_format_residual_frame_packed   residual text of one frame (same bytes as so_format_residual_frame); INT64_MIN when
 *                                     the stream is not a valid sequence of lists for these split flags
 *   so_symbols_to_levels              inverse RLE of whole frames -> levels i16 [n_frames][height][width], frames in
 *                                     parallel on host threads; SO_E_INVALID on a corrupt stream
 *   so_write_bitstream_files_symbols  so_write_bitstream_files with the residual text formatted from the symbols */
int64_t so_format_residual_frame_packed(const uint8_t* split, const int16_t* symbols, int64_t n_symbols, int n_blocks, int block_size,
                                        char* dst, int64_t cap);
int so_symbols_to_levels(const uint8_t* split, const int16_t* symbols, const uint64_t* sym_pos, const uint32_t* sym_count, int n_frames,
                         int width, int height, int block_size, int16_t* levels, int n_threads);
int so_write_bitstream_files_symbols(const uint8_t* frame_types, const uint8_t* split, const int16_t* mv, const int16_t* symbols,
                                     const uint64_t* sym_pos, const uint32_t* sym_count, const int32_t* qp_rows_per_frame, int n_frames,
                                     int width, int height, int block_size, const char* mv_path, const char* residual_path, int n_threads);

/* Decoder (decoder.py:487-545 `decode` with decode_frame_inter :97 / decode_frame_intra :330) on packed arrays, host
 * buffers in and out, one sequence (unit).  frame_types u8 [n_frames]; split / mv / levels as so_encode_sequence writes
 * them; qp_rows_per_frame i32 [n_frames][height/block_size] or NULL (RCFlag off).  reset_at_intra != 0 clears the
 * reference list at I frames like decoder.py:520 does; 0 keeps the ENCODER's list semantics (Encoder.py:1864-1867),
 * which is what round-trips streams encoded with nRefFrames > 1 (quirk Q7); in that mode the I frames a scene cut puts into
 * a ParallelMode-1 stream are decoded as intra (decoder.py:504-509 decodes every frame of such a stream as inter).
 * out_frames u8 [n_frames][height][width]. */
int so_decode_sequence(so_ctx* ctx, const uint8_t* frame_types, const uint8_t* split, const int16_t* mv, const int16_t* levels,
                       const int32_t* qp_rows_per_frame, int n_frames, int reset_at_intra, uint8_t* out_frames);

/* Timing of the last so_seq_run, CUDA events on the context stream (ms): [0] whole device region, [1] motion-search
 * kernels only (exhaustive search: the me_ring_kernel / me_tma_kernel launches; fast ME: the chain kernel; intra frames: intra search),
 * [2] transform/quant/recon kernels, [3] number of kernel launches.  Waits for the run to finish.  [1] and [2] cover the
 * frames that carry per-kernel events (see so_last_search_timing). */
int so_last_timing(so_ctx* ctx, double out[4]);
int so_last_me_launches(so_ctx* ctx);    /* number of launches covered by timing [1] */
/* Per-kernel CUDA events are recorded on every 8th frame of a sequence only (they serialise the stream: ~10 us per frame;
 * SO_TIMING_STRIDE=n in the environment changes the stride), so timing [1] and [2] above are sums over those frames.
 * The exhaustive-search kernel alone (me_ring_kernel / me_tma_kernel) in the last so_seq_run: out[0] = summed event time of
 * its timed launches (ms), out[1] = number of timed launches, out[2] = frames with per-kernel events, out[3] = frames of
 * the run.  out[0] / out[1] is the per-launch duration bench.py's roofline uses. */
int so_last_search_timing(so_ctx* ctx, double out[4]);
/* The inter finish kernel (transform / quantisation / RLE size / reconstruction of P frames, Encoder.py:779-827, :1086) alone:
 * out[0] = summed event time of its timed launches (ms), out[1] = timed launches (all units of a frame are one launch),
 * out[2] = the sum over all transform kernels incl. intra frames (= so_last_timing [2]), out[3] = frames with events. */
int so_last_finish_timing(so_ctx* ctx, double out[4]);
/* Counters of the SO_FLAG_SEA search since the context was created (waits for the device): out[0] = exact SADs computed
 * (predictor candidates + candidates that passed the bound; the plain search computes one per valid candidate), out[1] = launches
 * (one per P frame, all units), out[2] = out[3] = 0 (reserved). */
int so_sea_stats(so_ctx* ctx, uint64_t out[4]);

/* Host-side text formatters, byte-identical to the reference's (Encoder.py:1419-1542 with canonical integers).
 * Return the number of bytes written (excluding the terminating NUL), or the required size (negative) when cap
 * is too small.  qp_rows may be NULL (RCFlag off). */
int64_t so_format_mv_frame(int frame_type, const uint8_t* split, const int16_t* mv, int n_blocks, int blocks_per_row,
                           const int32_t* qp_rows, char* dst, int64_t cap);
int64_t so_format_residual_frame(const uint8_t* split, const int16_t* levels, int width, int height, int block_size,
                                 char* dst, int64_t cap);

/* transmit_bitstream (Encoder.py:1544-1573) for a whole sequence: both text files, frames formatted in parallel on host
 * threads (n_threads <= 0: hardware concurrency) and written in order; byte-identical to joining the per-frame formatters
 * with '\n'.  qp_rows_per_frame i32 [n_frames][height / block_size] or NULL.  Host only: no device needed. */
int so_write_bitstream_files(const uint8_t* frame_types, const uint8_t* split, const int16_t* mv, const int16_t* levels,
                             const int32_t* qp_rows_per_frame, int n_frames, int width, int height, int block_size,
                             const char* mv_path, const char* residual_path, int n_threads);

/* decode_differential_entropy (decoder.py:590-690): the two text streams -> packed arrays (as so_encode_sequence writes
 * them), lines parsed in parallel on host threads.  frame_types u8 [n_frames], split u8 [n_frames][n_blocks], mv i16
 * [n_frames][n_blocks][4][3], levels i16 [n_frames][height][width], qp_rows i32 [n_frames][height / block_size] (filled when
 * rc_on != 0).  SO_E_INVALID on unreadable, short or malformed input.  Host only: no device needed. */
int so_parse_bitstream_files(const char* mv_path, const char* residual_path, int n_frames, int width, int height, int block_size,
                             int rc_on, uint8_t* frame_types, uint8_t* split, int16_t* mv, int16_t* levels, int32_t* qp_rows,
                             int n_threads);

#ifdef __cplusplus
}
#endif
#endif /* STREAMOPTIMA_B200_H */
