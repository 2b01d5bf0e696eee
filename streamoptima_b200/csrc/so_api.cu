// C ABI + host-side orchestration of the StreamOptima B200 encode path (see include/streamoptima_b200.h).
// Host logic restates the frame loop of Encoder.py:1790-1898 and the two per-frame flows (:1582, :1644).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "so_kernels.cuh"
#include "so_me_tma.cuh"
#include "so_me_ring.cuh"
#include "so_me_ring2.cuh"
#include "so_me_sea.cuh"
#include <cstdlib>
#include <fcntl.h>
#include <unistd.h>
#include <algorithm>
#include <climits>

#define CU(expr)                                                                                  \
    do {                                                                                          \
        cudaError_t e__ = (expr);                                                                 \
        if (e__ != cudaSuccess) {                                                                 \
            set_err(ctx, std::string(#expr) + ": " + cudaGetErrorString(e__));                    \
            return SO_E_CUDA;                                                                     \
        }                                                                                         \
    } while (0)

static_assert(sizeof(so_params) == 64 && sizeof(so_frame_stats) == 32, "ABI struct layout (mirrored in _native.py)");
static std::string g_create_err;

// Launch with programmatic dependent launch (PDL): the grid may be scheduled while its predecessor in the stream is still
// draining; the kernel itself waits (griddepcontrol.wait) before it touches anything the predecessor wrote.
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    static const bool no_pdl = std::getenv("SO_NO_PDL") != nullptr;      // A/B switch
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = no_pdl ? 0 : 1;
    return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

struct so_ctx {
    so_params p{};
    int device = 0;
    cudaStream_t stream = nullptr;          // used by so_encode_sequence
    std::string err;
    FrameGeom g{};
    int nblk = 0, batch = 1;
    size_t frame_px = 0;
    // reference ring: [unit][slot][phase 4][shift 4][H][pitch]
    uint8_t* ring = nullptr;
    size_t plane_bytes = 0, slot_stride = 0, unit_stride = 0;
    int nslots = 0;
    // Reference-list state.  Whole-sequence calls advance all units in lock step (one state, rstates[0], cur_unit = -1); the
    // per-frame calls with unit >= 0 give every unit its own chain (rstates[unit]).
    struct RingState {
        std::vector<int> list;              // ring slots in list order (oldest first)
        std::vector<char> slot_u8;          // slot holds a uint8 reconstruction (false: the float 128 frame)
        std::vector<int> slot_wrap;         // wrap mode the half-pel planes of the slot were built with (-1 none)
        std::vector<char> slot_sea;         // the slot's packed quadrant bytes (successive elimination) match its planes
        bool operator==(const RingState& o) const { return list == o.list && slot_u8 == o.slot_u8 && slot_wrap == o.slot_wrap; }
    };
    std::vector<RingState> rstates;
    bool lockstep = true;
    int cur_unit = -1;                      // unit the ring helpers act on (-1: all units, lock step)
    RingState& rs() { return rstates[cur_unit < 0 ? 0 : cur_unit]; }
    int u0() const { return cur_unit < 0 ? 0 : cur_unit; }              // first unit / number of units of a ring operation
    int un() const { return cur_unit < 0 ? batch : 1; }
    // scratch
    MeResult *me_parent = nullptr, *me_sub = nullptr;       // exhaustive search: packed keys (all ones between frames); fast ME: records
    MeResult *in_parent = nullptr, *in_sub = nullptr;       // intra search results (separate: they must not disturb the keys)
    int16_t* res_frame = nullptr;
    int32_t* band = nullptr;
    int* qp_rows_dev = nullptr;
    std::vector<int> qp_rows;
    unsigned int* me_work = nullptr;        // chunk counter of the item-ring search kernel
    int me_key_fmt = 1;                     // format of the packed keys the last exhaustive search left (FlowArgs::me_packed)
    // successive elimination (so_me_sea.cuh, SO_FLAG_SEA): packed quadrant bytes [unit][slot][phase][H][W], last winners, counters
    uint32_t *sea_pq = nullptr, *sea_prev = nullptr;
    unsigned int* sea_ctr = nullptr;
    // SO_FLAG_SEA_AUTO: mapped host words {exact SADs of the last finished pruned launch, its sequence number}, what the host
    // has seen of them, candidates-per-launch bookkeeping and the number of frames the plain kernel still runs
    unsigned int *sea_host = nullptr, *sea_host_dev = nullptr;
    unsigned int sea_seq = 0, sea_seen = 0;
    double sea_items[1024] = {};            // candidates of the launches in flight (the host runs up to a sequence ahead), by sequence number
    int sea_cooldown = 0, sea_skip = 2;     // the first pruned frames of a context start without predictors: not judged
    double sea_avg = -1.0;                  // running average of exact SADs per (block, reference, phase plane)
    uint8_t* fm_table = nullptr;            // fast ME: per-block transition tables around the previous frame's predictors
    short4* fm_state = nullptr;             // fast ME: predictor (x, y, ref) every block used in the last P frame, [batch][nblk]
    int2 *fs_F = nullptr, *fs_entry = nullptr;   // fast ME scan: per-chunk composed transition functions / chunk entry predictors
    int fs_L = 0, fs_nchunks = 0;
    int* qp_blocks_dev = nullptr;           // ROI extension: [frames][nblk] per-block QPs of the next sequence, or nullptr
    int qp_blocks_frames = 0;
    // sequence buffers
    uint8_t *sq_frames = nullptr, *sq_split = nullptr, *sq_recon = nullptr;
    int16_t *sq_mv = nullptr, *sq_levels = nullptr;
    uint32_t *sq_rows = nullptr, *sq_blklen = nullptr;      // sq_blklen: RLE symbols per block, [unit][frame][nblk] (count pass of the symbol packer)
    so_frame_stats* sq_stats = nullptr;
    size_t sq_cap_frames = 0;               // capacity in (unit*frame) frames
    // run-level symbol streams of the resident sequence (so_seq_symbols)
    uint32_t *sym_lens = nullptr, *sym_offs = nullptr, *sym_tot = nullptr, *sym_boffs = nullptr;
    size_t sym_sub_cap_frames = 0;          // capacity of sym_lens / sym_offs (per-sub-block API, so_seq_symbols)
    uint32_t* cur_blk_len = nullptr;        // block symbol counts of the frame being encoded (sequence path), or nullptr
    int16_t* sym_data = nullptr;
    size_t sym_cap_frames = 0, sym_frame_stride = 0;
    uint32_t* h_tot = nullptr;              // pinned: symbols per frame, [unit][frame]
    bool sym_resident = false;              // sym_data holds the symbols of the whole resident sequence
    so_symbol_out* sym_out = nullptr;       // armed by so_set_symbol_output: sequence encodes deliver packed symbols
    uint64_t sym_used = 0;                  // symbols handed out so far in the running sequence encode
    bool sym_overflow = false;
    double timing[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    bool timing_pending = false;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    int sq_units = 0, sq_nframes = 0;
    // chunked copy/compute overlap of so_encode_sequence: H2D of chunk c+1 and D2H of chunk c-1 run on their own streams
    // while chunk c is encoded
    struct Pipe {
        bool active = false;
        int chunk = 8, nchunks = 0;
        const uint8_t* h_frames = nullptr;
        uint8_t *h_split = nullptr, *h_recon = nullptr;
        int16_t *h_mv = nullptr, *h_levels = nullptr;
        uint32_t* h_rows = nullptr;
        so_frame_stats* h_stats = nullptr;
        std::vector<cudaEvent_t> up, done, tot;
        // file ingest (so_encode_yuv420_file): luma planes are pread() chunk by chunk into three rotating pinned buffers
        int fd = -1;
        int src_w = 0, src_h = 0, first_frame = 0;
        uint8_t* stage[3] = {nullptr, nullptr, nullptr};
        size_t stage_bytes = 0;
    } pipe;
    cudaStream_t st_h2d = nullptr, st_d2h = nullptr;
    long launches = 0;
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used = 0;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev_me, ev_tq;
    std::vector<size_t> ev_xs;              // indices into ev_me of the exhaustive-search kernel launches
    std::vector<size_t> ev_tq_inter;        // indices into ev_tq of the inter finish kernel launches (the rest are intra frames)
    bool timing_on = false;
    int timed_frames = 0, total_frames = 0;    // frames of the last so_seq_run with / without per-kernel events
    const int* cur_qp_blocks = nullptr;     // per-block QPs of the frame being encoded (ROI extension)
    bool stats_prezeroed = false;           // so_seq_run zeroes the statistics of the whole sequence with one memset
};

static void set_err(so_ctx* ctx, const std::string& s) {
    if (ctx) ctx->err = s; else g_create_err = s;
}

extern "C" int so_abi_version(void) { return SO_ABI_VERSION; }
extern "C" const char* so_last_error(const so_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

extern "C" int so_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return SO_E_CUDA; }
    int ok = 0;
    for (int d = 0; d < n; ++d) {
        int major = 0;
        cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d);
        if (major == 10) ++ok;
    }
    return ok;
}

// ---------------------------------------------------------------------------------------------------------
// tables
// ---------------------------------------------------------------------------------------------------------
static int upload_tables(so_ctx* ctx) {
    uint16_t pos[340];
    const int sizes[4] = {2, 4, 8, 16};
    const int offs[4] = {0, 4, 20, 84};
    for (int s = 0; s < 4; ++s) {
        const int N = sizes[s];
        int p = 0;
        for (int d = 0; d < 2 * N - 1; ++d) {
            int i = d < N ? 0 : d - N + 1, j = d < N ? d : N - 1;
            while (i < N && j >= 0) { pos[offs[s] + i * N + j] = (uint16_t)p++; ++i; --j; }
        }
    }
    CU(cudaMemcpyToSymbol(c_scanpos, pos, sizeof(pos)));
    return SO_OK;
}

// ---------------------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------------------
static void free_seq(so_ctx* c) {
    cudaFree(c->sq_frames); cudaFree(c->sq_split); cudaFree(c->sq_recon); cudaFree(c->sq_mv);
    cudaFree(c->sq_levels); cudaFree(c->sq_rows); cudaFree(c->sq_stats); cudaFree(c->sq_blklen);
    c->sq_frames = c->sq_split = c->sq_recon = nullptr; c->sq_mv = c->sq_levels = nullptr;
    c->sq_rows = c->sq_blklen = nullptr; c->sq_stats = nullptr; c->sq_cap_frames = 0;
    cudaFree(c->sym_lens); cudaFree(c->sym_offs); cudaFree(c->sym_data); cudaFree(c->sym_tot); cudaFree(c->sym_boffs);
    if (c->h_tot) cudaFreeHost(c->h_tot);
    c->sym_lens = c->sym_offs = c->sym_tot = c->sym_boffs = nullptr; c->sym_data = nullptr; c->h_tot = nullptr;
    c->sym_cap_frames = c->sym_sub_cap_frames = 0;
    c->sym_resident = false;
}

extern "C" void so_ctx_destroy(so_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    free_seq(c);
    cudaFree(c->ring); cudaFree(c->me_parent); cudaFree(c->me_sub); cudaFree(c->in_parent); cudaFree(c->in_sub);
    cudaFree(c->res_frame); cudaFree(c->band);
    cudaFree(c->qp_rows_dev); cudaFree(c->qp_blocks_dev); cudaFree(c->me_work); cudaFree(c->sea_pq); cudaFree(c->sea_prev); cudaFree(c->sea_ctr);
    if (c->sea_host) cudaFreeHost(c->sea_host);
    cudaFree(c->fm_table); cudaFree(c->fm_state);
    cudaFree(c->fs_F); cudaFree(c->fs_entry);
    for (auto e : c->ev_pool) cudaEventDestroy(e);
    for (auto b : c->pipe.stage) if (b) cudaFreeHost(b);
    for (auto e : c->pipe.up) cudaEventDestroy(e);
    for (auto e : c->pipe.done) cudaEventDestroy(e);
    for (auto e : c->pipe.tot) cudaEventDestroy(e);
    if (c->st_h2d) cudaStreamDestroy(c->st_h2d);
    if (c->st_d2h) cudaStreamDestroy(c->st_d2h);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

extern "C" int so_ctx_create(so_ctx** out, const so_params* p, int device) {
    so_ctx* ctx = nullptr;
    if (!out || !p) { set_err(nullptr, "null argument"); return SO_E_INVALID; }
    *out = nullptr;
    auto bad = [&](const char* m) { set_err(nullptr, m); return SO_E_INVALID; };
    if (p->block_size != 4 && p->block_size != 8 && p->block_size != 16) return bad("block_size must be 4, 8 or 16");
    if (p->width <= 0 || p->height <= 0 || p->width % p->block_size || p->height % p->block_size)
        return bad("width/height must be positive multiples of block_size (Encoder.py:1382)");
    if (p->search_range < 0 || p->search_range > 63) return bad("search_range must be 0..63");
    if (p->n_ref_frames < 1 || p->n_ref_frames > SO_MAX_REF) return bad("n_ref_frames must be 1..8");
    if (p->qp < 0 || p->qp > 15) return bad("qp out of range");
    if (p->parallel_mode < 0 || p->parallel_mode > 2) return bad("parallel_mode must be 0, 1 or 2 (3 is broken in the reference)");
    if (p->rc_flag < 0 || p->rc_flag > 2) return bad("rc_flag must be 0, 1 or 2");
    if ((p->flags & SO_FLAG_VBS) && (p->flags & SO_FLAG_FAST_ME) && p->parallel_mode != 0)
        return bad("VBS + fast_me in ParallelMode 1/2 raises UnboundLocalError in the reference (Encoder.py:616)");
    if (p->intra_dur < 1) return bad("intra_dur must be >= 1");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) {
        cudaGetLastError();
        set_err(nullptr, "no usable CUDA device (this library has no CPU fallback)");
        return SO_E_CUDA;
    }
    int major = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
    if (major != 10) { set_err(nullptr, "device is not sm_100 (kernels are built for sm_100a only)"); return SO_E_CUDA; }

    ctx = new so_ctx();
    ctx->p = *p;
    ctx->device = device;
    ctx->batch = p->max_batch > 0 ? p->max_batch : 1;
    FrameGeom& g = ctx->g;
    g.W = p->width; g.H = p->height; g.bs = p->block_size;
    g.pitch = (g.W + 15) / 16 * 16;
    g.nbx = g.W / g.bs; g.nby = g.H / g.bs;
    g.r = p->search_range;
    g.fme = (p->flags & SO_FLAG_FME) ? 1 : 0;
    g.R = g.fme ? 2 * g.r : g.r;
    g.nref = 1;
    ctx->nblk = g.nbx * g.nby;
    ctx->frame_px = (size_t)g.W * g.H;
    auto fail = [&](int code) { g_create_err = ctx->err; so_ctx_destroy(ctx); return code; };
#define CUC(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { ctx->err = std::string(#expr) + ": " + cudaGetErrorString(e__); return fail(SO_E_CUDA); } } while (0)
    CUC(cudaSetDevice(device));
    CUC(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    CUC(cudaStreamCreateWithFlags(&ctx->st_h2d, cudaStreamNonBlocking));
    CUC(cudaStreamCreateWithFlags(&ctx->st_d2h, cudaStreamNonBlocking));
    ctx->nslots = p->n_ref_frames;
    ctx->plane_bytes = (size_t)g.pitch * g.H;
    ctx->slot_stride = ctx->plane_bytes * 16;
    ctx->unit_stride = ctx->slot_stride * ctx->nslots;
    CUC(cudaMalloc(&ctx->ring, ctx->unit_stride * ctx->batch));
    CUC(cudaMemset(ctx->ring, 0, ctx->unit_stride * ctx->batch));
    CUC(cudaMalloc(&ctx->me_parent, sizeof(MeResult) * ctx->nblk * ctx->batch));
    CUC(cudaMalloc(&ctx->me_sub, sizeof(MeResult) * ctx->nblk * 4 * ctx->batch));
    CUC(cudaMalloc(&ctx->in_parent, sizeof(MeResult) * ctx->nblk * ctx->batch));
    CUC(cudaMalloc(&ctx->in_sub, sizeof(MeResult) * ctx->nblk * 4 * ctx->batch));
    CUC(cudaMalloc(&ctx->res_frame, sizeof(int16_t) * ctx->frame_px * ctx->batch));
    CUC(cudaMalloc(&ctx->band, sizeof(int32_t) * ctx->frame_px * ctx->batch));
    CUC(cudaMalloc(&ctx->qp_rows_dev, sizeof(int) * g.nby));
    ctx->rstates.resize(ctx->batch);
    for (auto& r : ctx->rstates) { r.slot_u8.assign(ctx->nslots, 0); r.slot_wrap.assign(ctx->nslots, -1); r.slot_sea.assign(ctx->nslots, 0); }
    if (upload_tables(ctx) != SO_OK) return fail(SO_E_CUDA);
    *out = ctx;
    return SO_OK;
}

// set_Qp for the whole sequence (Encoder.py:948): changes const_init_Qp of subsequent encodes
extern "C" int so_set_qp(so_ctx* ctx, int qp) {
    if (!ctx) return SO_E_INVALID;
    if (qp < 0 || qp > 15) { set_err(ctx, "qp out of range"); return SO_E_INVALID; }
    ctx->p.qp = qp;
    return SO_OK;
}

// ROI extension (no counterpart in the reference, whose bitstream carries a QP only at block-row starts): per-block QPs
// for the final quantisation of the next sequence(s), qp_blocks i32 [n_frames][n_blocks] (shared by all units); NULL clears.
extern "C" int so_set_block_qps(so_ctx* ctx, const int32_t* qp_blocks, int n_frames) {
    if (!ctx) return SO_E_INVALID;
    CU(cudaSetDevice(ctx->device));
    cudaFree(ctx->qp_blocks_dev);
    ctx->qp_blocks_dev = nullptr; ctx->qp_blocks_frames = 0;
    if (!qp_blocks || n_frames < 1) return SO_OK;
    const size_t n = (size_t)n_frames * ctx->nblk;
    for (size_t i = 0; i < n; ++i)
        if (qp_blocks[i] < 0 || qp_blocks[i] > 15) { set_err(ctx, "block qp out of range"); return SO_E_INVALID; }
    CU(cudaMalloc(&ctx->qp_blocks_dev, n * sizeof(int)));
    CU(cudaMemcpy(ctx->qp_blocks_dev, qp_blocks, n * sizeof(int), cudaMemcpyHostToDevice));
    ctx->qp_blocks_frames = n_frames;
    return SO_OK;
}

extern "C" int so_set_row_qps(so_ctx* ctx, const int32_t* qp_rows, int n) {
    if (!ctx) return SO_E_INVALID;
    if (!qp_rows || n != ctx->g.nby) { set_err(ctx, "qp_rows must have height/block_size entries"); return SO_E_INVALID; }
    for (int i = 0; i < n; ++i)
        if (qp_rows[i] < 0 || qp_rows[i] > 15) { set_err(ctx, "row qp out of range"); return SO_E_INVALID; }
    CU(cudaSetDevice(ctx->device));
    ctx->qp_rows.assign(qp_rows, qp_rows + n);
    CU(cudaMemcpy(ctx->qp_rows_dev, qp_rows, sizeof(int) * n, cudaMemcpyHostToDevice));
    return SO_OK;
}

// ---------------------------------------------------------------------------------------------------------
// reference ring
// ---------------------------------------------------------------------------------------------------------
static uint8_t* slot_ptr(so_ctx* c, int slot) { return c->ring + (size_t)slot * c->slot_stride; }

static int take_free_slot(so_ctx* c) {
    for (int s = 0; s < c->nslots; ++s) {
        bool used = false;
        for (int l : c->rs().list) used = used || (l == s);
        if (!used) return s;
    }
    return -1;
}

// Choose the chain the following ring operations act on: SO_ALL_UNITS (-1) = every unit in lock step, otherwise one unit.
// Going from lock step to single units forks the shared state; going back needs identical chains (or a reset).
static int select_unit(so_ctx* ctx, int unit, bool resetting = false) {
    if (unit < -1 || unit >= ctx->batch) { set_err(ctx, "unit out of range (0 .. max_batch-1, or SO_ALL_UNITS)"); return SO_E_INVALID; }
    if (unit >= 0) {
        if (ctx->lockstep) { for (int u = 1; u < ctx->batch; ++u) ctx->rstates[u] = ctx->rstates[0]; ctx->lockstep = false; }
    } else if (!ctx->lockstep) {
        if (!resetting) {
            for (int u = 1; u < ctx->batch; ++u)
                if (!(ctx->rstates[u] == ctx->rstates[0])) {
                    set_err(ctx, "SO_ALL_UNITS after per-unit calls left the chains in different states: so_ref_reset(ctx, SO_ALL_UNITS) first");
                    return SO_E_STATE;
                }
        }
        ctx->lockstep = true;
    }
    ctx->cur_unit = unit;
    return SO_OK;
}

// exhaustive-search key arrays := all ones (the identity of the atomicMin merge)
static int keys_reset(so_ctx* ctx, cudaStream_t st) {
    const size_t u0 = ctx->u0(), un = ctx->un();
    CU(cudaMemsetAsync(ctx->me_parent + u0 * ctx->nblk, 0xFF, sizeof(MeResult) * ctx->nblk * un, st));
    CU(cudaMemsetAsync(ctx->me_sub + u0 * ctx->nblk * 4, 0xFF, sizeof(MeResult) * ctx->nblk * 4 * un, st));
    return SO_OK;
}

static int ref_reset_impl(so_ctx* ctx, cudaStream_t st, bool reset_keys);

extern "C" int so_ref_reset(so_ctx* ctx, int unit, void* stream) {
    if (!ctx) return SO_E_INVALID;
    CU(cudaSetDevice(ctx->device));
    int rc = select_unit(ctx, unit, true);
    if (rc) return rc;
    return ref_reset_impl(ctx, (cudaStream_t)stream, true);
}

static int ref_reset_impl(so_ctx* ctx, cudaStream_t st, bool reset_keys) {
    so_ctx::RingState& R = ctx->rs();
    R.list.clear();
    if (reset_keys) {
        const int rc = keys_reset(ctx, st);
        if (rc) return rc;
    }
    const int s = 0;
    // ref_frames = [np.ones((h, w)) * 128]  (Encoder.py:1798): a float frame -> never triggers the uint8 wrap
    const size_t n16 = ctx->slot_stride / 16;
    ring_fill_kernel<<<dim3((unsigned)((n16 + 255) / 256), ctx->un()), 256, 0, st>>>(slot_ptr(ctx, s) + (size_t)ctx->u0() * ctx->unit_stride,
                                                                                     ctx->unit_stride, n16, 0x80808080u);
    ctx->launches++;
    CU(cudaGetLastError());
    R.list.push_back(s);
    R.slot_u8[s] = 0;
    R.slot_wrap[s] = 0;          // planes of a constant frame are the constant in both modes
    R.slot_sea[s] = 0;
    return SO_OK;
}

// FIFO append (Encoder.py:1864-1867) to the selected chain(s); recon_dev is [units][H][W] dense with unit stride src_unit_stride
static int ring_push(so_ctx* ctx, const uint8_t* recon_dev, size_t src_unit_stride, int units, cudaStream_t st) {
    so_ctx::RingState& R = ctx->rs();
    if ((int)R.list.size() >= ctx->p.n_ref_frames) R.list.erase(R.list.begin());
    const int s = take_free_slot(ctx);
    if (s < 0) { set_err(ctx, "reference ring has no free slot"); return SO_E_STATE; }
    const FrameGeom& g = ctx->g;
    R.list.push_back(s);
    R.slot_u8[s] = 1;
    // One kernel stores the frame and derives its half-pel phases and byte-shifted copies.  The uint8-wrap mode (quirk
    // Q1) of the next inter frame is already known: the list does not change before it (ensure_planes re-derives the
    // planes from the stored frame in the one case it does: ParallelMode 1 resets the list every frame).
    bool all_u8 = true;
    for (int l : R.list) all_u8 = all_u8 && R.slot_u8[l];
    const int wrap = (g.fme && all_u8) ? 1 : 0;
    if (!g.fme && g.W % 16 == 0 && reinterpret_cast<uintptr_t>(recon_dev) % 16 == 0 && src_unit_stride % 16 == 0)
        CU(launch_pdl(ring_shift_kernel, dim3((g.W / 16 + 127) / 128, g.H, units), dim3(128), 0, st,
                      slot_ptr(ctx, s) + (size_t)ctx->u0() * ctx->unit_stride, ctx->unit_stride,
                      ctx->plane_bytes, recon_dev, src_unit_stride, g.W, g.W, g.pitch));
    else
    CU(launch_pdl(ring_planes_kernel, dim3((g.W / 4 + 127) / 128, g.H, units), dim3(128), 0, st,
                  slot_ptr(ctx, s) + (size_t)ctx->u0() * ctx->unit_stride, ctx->unit_stride,
                  ctx->plane_bytes, recon_dev, src_unit_stride, g.W, g.W, g.H, g.pitch, g.fme, wrap, 1));
    ctx->launches++;
    CU(cudaGetLastError());
    R.slot_wrap[s] = wrap;
    R.slot_sea[s] = 0;
    return SO_OK;
}

extern "C" int so_ref_push(so_ctx* ctx, int unit, const uint8_t* recon_dev, void* stream) {
    if (!ctx || !recon_dev) return SO_E_INVALID;
    CU(cudaSetDevice(ctx->device));
    int rc = select_unit(ctx, unit);
    if (rc) return rc;
    return ring_push(ctx, recon_dev, ctx->frame_px, ctx->un(), (cudaStream_t)stream);
}

static RefRing make_ring(so_ctx* c) {
    RefRing r;
    r.base = c->ring; r.unit_stride = c->unit_stride; r.slot_stride = c->slot_stride; r.plane_stride = c->plane_bytes * 4;
    const std::vector<int>& list = c->rs().list;
    for (int i = 0; i < SO_MAX_REF; ++i) r.slot[i] = i < (int)list.size() ? list[i] : 0;
    return r;
}

static void ev_pair(so_ctx* ctx, std::vector<std::pair<cudaEvent_t, cudaEvent_t>>& v, cudaStream_t st, bool start) {
    if (!ctx->timing_on) return;
    if (start) {
        if (ctx->ev_used + 2 > ctx->ev_pool.size()) {
            for (int i = 0; i < 64; ++i) {
                cudaEvent_t e;
                if (cudaEventCreate(&e) != cudaSuccess) { cudaGetLastError(); ctx->timing_on = false; return; }   // no events: this frame goes untimed
                ctx->ev_pool.push_back(e);
            }
        }
        cudaEvent_t a = ctx->ev_pool[ctx->ev_used++], b = ctx->ev_pool[ctx->ev_used++];
        v.emplace_back(a, b);
        cudaEventRecord(a, st);
    } else {
        cudaEventRecord(v.back().second, st);
    }
}

// ---------------------------------------------------------------------------------------------------------
// kernel dispatch
// ---------------------------------------------------------------------------------------------------------
// ---- TMA: one 3-D tensor map {W, H, planes} over the whole reference ring, box = one search window --------------
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_tmapEncodeTiled tmap_encoder(so_ctx* ctx) {
    static PFN_tmapEncodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p ||
            q != cudaDriverEntryPointSuccess) {
            set_err(ctx, "cuTensorMapEncodeTiled not available");
            return nullptr;
        }
        fn = (PFN_tmapEncodeTiled)p;
    }
    return fn;
}

static int make_map3d(so_ctx* ctx, const void* base, cuuint64_t w, cuuint64_t h, cuuint64_t z, cuuint64_t row_stride, cuuint64_t z_stride,
                      int box_w, int box_h, CUtensorMap* map, int box_z = 1) {
    PFN_tmapEncodeTiled fn = tmap_encoder(ctx);
    if (!fn) return SO_E_CUDA;
    cuuint64_t gdim[3] = {w, h, z};
    cuuint64_t gstr[2] = {row_stride, z_stride};
    cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, (cuuint32_t)box_z};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_err(ctx, "cuTensorMapEncodeTiled failed with code " + std::to_string((int)r)); return SO_E_CUDA; }
    return SO_OK;
}

// one 3-D tensor map {W, H, planes} over the whole reference ring, box = one search window
static int make_ring_map(so_ctx* ctx, int box_w, int box_h, CUtensorMap* map, int box_z = 1) {
    const FrameGeom& g = ctx->g;
    return make_map3d(ctx, ctx->ring, g.W, g.H, (cuuint64_t)ctx->batch * ctx->nslots * 16, g.pitch, ctx->plane_bytes, box_w, box_h, map, box_z);
}

template <int BS, int NDX, int G, bool QUAD = false>
static cudaError_t launch_me_tma(const CUtensorMap& map, const CUtensorMap& cmap, const MeTmaArgs& a, int grid, int threads, size_t smem, cudaStream_t st) {
    static bool attr_done[64] = {};      // per device ordinal; beyond 64 devices the attribute is simply set again
    int dev = 0; cudaGetDevice(&dev);
    if (dev >= 64 || !attr_done[dev]) {
        cudaError_t e = cudaFuncSetAttribute(me_tma_kernel<BS, NDX, G, QUAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return e;
        if (dev < 64) attr_done[dev] = true;
    }
    me_tma_kernel<BS, NDX, G, QUAD><<<grid, threads, smem, st>>>(map, cmap, a);
    return cudaGetLastError();
}
template <int BS, int NDX>
static cudaError_t launch_me_tma_g(int G, const CUtensorMap& map, const CUtensorMap& cmap, const MeTmaArgs& a, int grid, int threads, size_t smem, cudaStream_t st) {
    return G == 3 ? launch_me_tma<BS, NDX, 3>(map, cmap, a, grid, threads, smem, st) : launch_me_tma<BS, NDX, 1>(map, cmap, a, grid, threads, smem, st);
}
template <int BS>
static cudaError_t launch_me_tma_n(int NDX, int G, const CUtensorMap& map, const CUtensorMap& cmap, const MeTmaArgs& a, int grid, int threads, size_t smem, cudaStream_t st) {
    switch (NDX) {
        case 1: return launch_me_tma_g<BS, 1>(G, map, cmap, a, grid, threads, smem, st);
        case 2: return launch_me_tma_g<BS, 2>(G, map, cmap, a, grid, threads, smem, st);
        case 3: return launch_me_tma_g<BS, 3>(G, map, cmap, a, grid, threads, smem, st);
        case 5: return launch_me_tma_g<BS, 5>(G, map, cmap, a, grid, threads, smem, st);
        default: return launch_me_tma_g<BS, 9>(G, map, cmap, a, grid, threads, smem, st);
    }
}

// ---- successive elimination (so_me_sea.cuh) -----------------------------------------------------------------------------------
static bool sea_enabled(const so_ctx* ctx) {
    static const char* env = std::getenv("SO_ME_SEA");             // tests / A-B: "1" forces it on, "0" off
    if (env && env[0] == '1') return true;
    if (env && env[0] == '0') return false;
    return (ctx->p.flags & SO_FLAG_SEA) != 0;
}

// SO_FLAG_SEA_AUTO: pruning pays while the pruned launches take few exact SADs per (block, reference, phase plane).  The count of
// the last FINISHED launch is read from mapped host memory without synchronising; when its running average exceeds the threshold
// the plain kernel runs for the next SEA_COOLDOWN P frames, then pruning is probed again (the first probe frame starts from stale
// predictors and is not judged).
static bool sea_pays(so_ctx* ctx) {
    if (!(ctx->p.flags & SO_FLAG_SEA_AUTO)) return true;
    constexpr double SEA_MAX_PER_ITEM = 7.0, SEA_EMA = 0.2;
    constexpr int SEA_COOLDOWN = 60;
    if (ctx->sea_host) {
        const unsigned int seq = reinterpret_cast<volatile unsigned int*>(ctx->sea_host)[1];
        if (seq != ctx->sea_seen) {
            const unsigned int evals = reinterpret_cast<volatile unsigned int*>(ctx->sea_host)[0];
            ctx->sea_seen = seq;
            const double items = ctx->sea_items[seq & 1023u];
            if (ctx->sea_skip > 0) --ctx->sea_skip;
            else if (ctx->sea_cooldown == 0 && items > 0) {
                // exponential average: one expensive frame (cold predictors after an I frame) does not switch the search
                const double x = std::min(evals / items, 3.0 * SEA_MAX_PER_ITEM);
                ctx->sea_avg = ctx->sea_avg < 0 ? x : (1.0 - SEA_EMA) * ctx->sea_avg + SEA_EMA * x;
                if (ctx->sea_avg > SEA_MAX_PER_ITEM) ctx->sea_cooldown = SEA_COOLDOWN;
            }
        }
    }
    if (ctx->sea_cooldown > 0) {
        if (--ctx->sea_cooldown == 0) ctx->sea_skip = 1;       // the probe frame that follows is not judged
        return false;
    }
    return true;
}

static int run_sea(so_ctx* ctx, const MeRingArgs& a, const uint8_t* cur, size_t cur_stride, int unit0, int units, cudaStream_t st) {
    const FrameGeom& g = ctx->g;
    const int nph = a.nph;
    if (!ctx->sea_pq) {
        const size_t nb = (size_t)ctx->nblk * ctx->batch;
        CU(cudaMalloc(&ctx->sea_pq, (size_t)ctx->batch * ctx->nslots * nph * ctx->frame_px * sizeof(uint32_t)));
        CU(cudaMalloc(&ctx->sea_prev, nb * sizeof(uint32_t)));
        CU(cudaMalloc(&ctx->sea_ctr, 256));
        CU(cudaMemsetAsync(ctx->sea_prev, 0xFF, nb * sizeof(uint32_t), st));
        CU(cudaMemsetAsync(ctx->sea_ctr, 0, 256, st));
        if (ctx->p.flags & SO_FLAG_SEA_AUTO) {
            CU(cudaHostAlloc(reinterpret_cast<void**>(&ctx->sea_host), 64, cudaHostAllocMapped));
            ctx->sea_host[0] = ctx->sea_host[1] = 0u;
            CU(cudaHostGetDevicePointer(reinterpret_cast<void**>(&ctx->sea_host_dev), ctx->sea_host, 0));
        }
    }
    SeaArgs s{};
    s.g = a.g;
    s.ring = ctx->ring; s.unit_stride = ctx->unit_stride; s.slot_stride = ctx->slot_stride; s.plane_bytes = ctx->plane_bytes;
    s.nph = nph;
    s.pq = ctx->sea_pq;
    s.pq_plane_stride = ctx->frame_px;
    s.pq_slot_stride = s.pq_plane_stride * (size_t)nph;
    s.pq_unit_stride = s.pq_slot_stride * (size_t)ctx->nslots;
    s.slot_packed = a.slot_packed;
    s.cur = cur + (size_t)unit0 * cur_stride; s.cur_unit_stride = cur_stride;
    s.out = a.out; s.out_unit_stride = a.out_unit_stride;
    s.prev = ctx->sea_prev; s.ctr = ctx->sea_ctr;
    s.host_stat = ctx->sea_host_dev;
    s.seq = ++ctx->sea_seq;
    ctx->sea_items[s.seq & 1023u] = (double)units * ctx->nblk * a.g.nref * nph;
    s.unit0 = unit0; s.units = units; s.nblk = ctx->nblk;
    // quadrant bytes of the references whose planes changed since they were last derived
    so_ctx::RingState& R = ctx->rs();
    for (int sl : R.list) {
        if (R.slot_sea[sl]) continue;
        CU(launch_pdl(sea_qplane_kernel, dim3((g.W + SQ_TX - 1) / SQ_TX, (g.H + SQ_TY - 1) / SQ_TY, units * nph), dim3(SQ_THREADS), 0, st,
                      (const uint8_t*)(slot_ptr(ctx, sl) + (size_t)unit0 * ctx->unit_stride), ctx->unit_stride, ctx->plane_bytes,
                      ctx->sea_pq + (size_t)unit0 * s.pq_unit_stride + (size_t)sl * s.pq_slot_stride, s.pq_unit_stride, s.pq_plane_stride,
                      g.W, g.H, g.pitch, nph));
        ctx->launches++;
        R.slot_sea[sl] = 1;
    }
    const int gpr = (s.g.nbx + SEA_NB - 1) / SEA_NB;
    // one warp per (reference, phase plane, block) triple in flight: 4 blocks x 1 plane (integer search, one reference) fill 4 warps
    const int sea_threads = s.g.nref * nph * SEA_NB <= 4 ? 128 : 256;
    CU(launch_pdl(sea_search_kernel, dim3(gpr * s.g.nby, units), dim3(sea_threads), 0, st, s));
    return SO_OK;
}

// counters of the successive-elimination search since the context was created: out[0] = exact SADs computed (predictors +
// candidates that passed the bound), out[1] = launches (one per P frame, all units), out[2], out[3] = 0 (reserved)
extern "C" int so_sea_stats(so_ctx* ctx, uint64_t* out4) {
    if (!ctx || !out4) return SO_E_INVALID;
    out4[0] = out4[1] = out4[2] = out4[3] = 0;
    if (!ctx->sea_ctr) return SO_OK;
    CU(cudaSetDevice(ctx->device));
    unsigned int h[4];
    CU(cudaMemcpy(h, ctx->sea_ctr, sizeof(h), cudaMemcpyDeviceToHost));
    out4[0] = (uint64_t)h[0] | ((uint64_t)h[1] << 32); out4[1] = h[2];
    return SO_OK;
}

// Item-ring search kernel (so_me_ring.cuh): 16x16 blocks, r = 16, DIRECT staging
static int run_me_ring(so_ctx* ctx, const uint8_t* cur, size_t cur_stride, int unit0, int units, MeResult* out, size_t out_stride,
                       cudaStream_t st, MeResult* out_sub = nullptr, size_t out_sub_stride = 0) {
    MeRingArgs a{};
    a.g = ctx->g;
    a.g.bs = 16; a.g.nbx = ctx->g.W / 16; a.g.nby = ctx->g.H / 16;
    a.g.nref = (int)ctx->rs().list.size();
    a.out = reinterpret_cast<unsigned long long*>(out + (size_t)unit0 * out_stride);
    a.out_unit_stride = out_stride;
    a.out_sub = out_sub ? reinterpret_cast<unsigned long long*>(out_sub + (size_t)unit0 * out_sub_stride) : nullptr;   // fused VBS search
    a.out_sub_unit_stride = out_sub_stride;
    a.units = units;
    a.nph = a.g.fme ? 4 : 1;
    a.items_per_unit = a.g.nbx * a.g.nby * a.g.nref * a.nph;
    a.z_per_unit = ctx->nslots * 16;
    a.z_unit0 = unit0 * a.z_per_unit;
    a.slot_packed = 0;
    for (int i = 0; i < SO_MAX_REF && i < (int)ctx->rs().list.size(); ++i) a.slot_packed |= (unsigned)(ctx->rs().list[i] & 15) << (4 * i);
    CUtensorMap map, cmap;
    int rc = make_ring_map(ctx, MR_WP, MR_BOXROWS, &map, 4);      // one box = the four shift planes of a phase
    if (rc) return rc;
    rc = make_map3d(ctx, cur + (size_t)unit0 * cur_stride, ctx->g.W, ctx->g.H, units, ctx->g.W, cur_stride ? cur_stride : ctx->frame_px, 16, 16, &cmap);
    if (rc) return rc;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
    static bool attr_done[64] = {};
    if (ctx->device >= 64 || !attr_done[ctx->device]) {
        CU(cudaFuncSetAttribute(me_ring_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, MR_SMEM));
        CU(cudaFuncSetAttribute(me_ring_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, MR_SMEM));
        CU(cudaFuncSetAttribute(me_ring2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, MR2_SMEM));
        CU(cudaFuncSetAttribute(me_ring2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, MR2_SMEM));
        if (ctx->device < 64) attr_done[ctx->device] = true;
    }
    if (!ctx->me_work) {      // {chunk counter, finished CTAs}: zero at every launch -- the last CTA of a launch resets both
        CU(cudaMalloc(&ctx->me_work, 256));
        CU(cudaMemsetAsync(ctx->me_work, 0, 256, st));
    }
    a.work = ctx->me_work;
    static const bool ring_v1 = std::getenv("SO_ME_RING_V1") != nullptr;      // A/B switch and cross-check: the first item-ring kernel
    ev_pair(ctx, ctx->ev_me, st, true);
    if (ctx->timing_on) ctx->ev_xs.push_back(ctx->ev_me.size() - 1);
    cudaError_t e;
    if (ring_v1) {
        const long long total = (long long)units * a.items_per_unit;
        const long long nchunks = (total + MR_CHUNK - 1) / MR_CHUNK;
        const int grid = nchunks < sms ? (int)nchunks : sms;
        if (out_sub) e = launch_pdl(me_ring_kernel<true>, dim3(grid), dim3(384), MR_SMEM, st, map, cmap, a);
        else e = launch_pdl(me_ring_kernel<false>, dim3(grid), dim3(512), MR_SMEM, st, map, cmap, a);
    } else {
        if (!out_sub && sea_enabled(ctx) && sea_pays(ctx)) {
            // successive elimination (so_me_sea.cuh): predictors -> bound filter -> exact SADs of the survivors, one kernel
            const int rc2 = run_sea(ctx, a, cur, cur_stride, unit0, units, st);
            if (rc2) return rc2;
            e = cudaSuccess;
        } else {
            MeRing2Args a2{};
            a2.b = a;
            a2.npairs = a.g.nbx * a.g.nby * a.g.nref;
            a2.chunks_per_unit = a.g.fme ? ((a2.npairs + 3) / 4) * 2 : (a2.npairs + 7) / 8;
            const long long nchunks = (long long)units * a2.chunks_per_unit;
            const int grid = nchunks < sms ? (int)nchunks : sms;
            ctx->me_key_fmt = 2;            // compact keys (me_get format 2)
            if (out_sub) e = launch_pdl(me_ring2_kernel<true>, dim3(grid), dim3(384), MR2_SMEM, st, map, cmap, a2);
            else {
                static const int nthr = std::getenv("SO_ME_RING_THREADS") ? atoi(std::getenv("SO_ME_RING_THREADS")) : 512;      // experiments: fewer search warps
                e = launch_pdl(me_ring2_kernel<false>, dim3(grid), dim3(nthr), MR2_SMEM, st, map, cmap, a2);
            }
        }
    }
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e == cudaSuccess && ctx->me_key_fmt == 2 && !out_sub && ctx->sea_prev && (ctx->p.flags & SO_FLAG_SEA_AUTO) && sea_enabled(ctx)) {
        // the plain kernel ran in auto mode: keep the pruned search's predictors fresh for the next probe
        SeaArgs sv{};
        sv.g = a.g; sv.out = a.out; sv.out_unit_stride = a.out_unit_stride; sv.prev = ctx->sea_prev;
        sv.unit0 = unit0; sv.units = units; sv.nblk = ctx->nblk;
        e = launch_pdl(sea_save_kernel, dim3((ctx->nblk + 255) / 256, units), dim3(256), 0, st, sv);
        ctx->launches++;
    }
    ev_pair(ctx, ctx->ev_me, st, false);
    if (e != cudaSuccess) { set_err(ctx, std::string("me_ring_kernel: ") + cudaGetErrorString(e)); return SO_E_CUDA; }
    ctx->launches++;
    return SO_OK;
}

// out_sub != nullptr asks for the fused VBS search (sub-block results from quadrant sums); *used_quad tells whether the
// geometry allowed it (16x16 blocks, DIRECT staging) -- otherwise the caller runs a second search on the sub-block grid
static int run_me_tma(so_ctx* ctx, const uint8_t* cur, size_t cur_stride, int unit0, int units, int bs, MeResult* out,
                      size_t out_stride, cudaStream_t st, MeResult* out_sub = nullptr, size_t out_sub_stride = 0, bool* used_quad = nullptr) {
    static const bool force_simple = std::getenv("SO_ME_SIMPLE") != nullptr;      // tests: cross-check of the packed kernels
    ctx->me_key_fmt = 1;
    if (bs < 4 || force_simple) {
        // 2x2 sub-blocks of VBS with block_size 4 (or the test switch): plain one-warp-per-block search
        if (used_quad) *used_quad = false;
        FlowArgs f{};
        f.g = ctx->g; f.g.nref = (int)ctx->rs().list.size();
        f.unit0 = unit0;
        f.cur = cur; f.cur_unit_stride = cur_stride;
        f.ring = make_ring(ctx);
        const int nb = (ctx->g.W / bs) * (ctx->g.H / bs);
        ev_pair(ctx, ctx->ev_me, st, true);
        if (ctx->timing_on) ctx->ev_xs.push_back(ctx->ev_me.size() - 1);
        me_simple_kernel<<<dim3((nb + 3) / 4, units), 128, 0, st>>>(f, bs, out, out_stride);
        ev_pair(ctx, ctx->ev_me, st, false);
        ctx->launches++;
        CU(cudaGetLastError());
        return SO_OK;
    }
    MeTmaArgs a{};
    a.g = ctx->g;
    a.g.bs = bs; a.g.nbx = ctx->g.W / bs; a.g.nby = ctx->g.H / bs;
    a.g.nref = (int)ctx->rs().list.size();
    a.cur = cur + (size_t)unit0 * cur_stride;
    a.cur_unit_stride = cur_stride;
    a.out = reinterpret_cast<unsigned long long*>(out + (size_t)unit0 * out_stride);
    a.out_unit_stride = out_stride;
    a.out_sub = nullptr; a.out_sub_unit_stride = 0;
    a.units = units;
    a.nph = a.g.fme ? 4 : 1;
    const int nb = a.g.nbx * a.g.nby;
    // ranges above 16 are tiled into chunks of 32 offsets per axis (the last chunk also takes offset +r): every chunk is
    // an r = 16 search with its own window origin, so the kernel geometry (9 x 3 candidates per task, 48-row windows) is shared
    const int r_real = a.g.r;
    const int r = r_real > 16 ? 16 : r_real;            // chunk half-range used for all geometry below
    a.cw = 2 * r;
    a.nxc = a.nyc = r_real > 16 ? (r_real + 15) / 16 : 1;
    a.items_per_unit = nb * a.g.nref * a.nph * a.nxc * a.nyc;
    const int ndx = (2 * r) / 4 + 1;
    const int avail[5] = {1, 2, 3, 5, 9};
    int NDX = 9;
    for (int v : avail) if (v >= ndx) { NDX = v; break; }
    const int G = ((2 * r + 1) % 3 == 0 || r >= 8) ? 3 : 1;
    a.NG = (2 * r + 1 + G - 1) / G;
    a.rows = bs + 2 * r;
    const int NW = NDX + bs / 4 - 1;
    int wp16 = (NW * 4 + 15) / 16;
    if (wp16 % 2 == 0) wp16 += 1;
    a.wpitch = wp16 * 16;
    const int tasks_per_item = 4 * a.NG;
    int SI = (352 + tasks_per_item - 1) / tasks_per_item;
    if (SI > 32) SI = 32;
    // DIRECT staging needs every window to start at a 16-byte aligned x (bs = 16, r multiple of 16), one group layout
    // without over-read (NG*G == 2r+1) and a TMA-addressable current frame (16-byte aligned rows)
    static const bool no_direct = std::getenv("SO_ME_NO_DIRECT") != nullptr;     // tests: force the EXPAND staging mode
    a.direct = (!no_direct && bs == 16 && r_real % 16 == 0 && r > 0 && a.NG * G == 2 * r + 1 && ctx->g.W % 16 == 0 &&
                (reinterpret_cast<uintptr_t>(cur) % 16 == 0) && (cur_stride % 16 == 0)) ? 1 : 0;
    static const bool no_ring = std::getenv("SO_ME_NO_RING") != nullptr;        // A/B switch: stage-based kernel of so_me_tma.cuh
    if (a.direct && r_real == 16 && !no_ring) {
        if (used_quad) *used_quad = out_sub != nullptr;
        return run_me_ring(ctx, cur, cur_stride, unit0, units, out, out_stride, st, out_sub, out_sub_stride);
    }
    size_t smem = 0;
    static const bool want_pad = std::getenv("SO_ME_NO_ROW_PAD") == nullptr;    // on by default; the switch is for A/B measurements
    a.row_pad = 0;
    if (a.direct) {
        a.nstage = 3;
        a.row_pad = want_pad ? 1 : 0;
        if (a.row_pad) a.nstage = 2;
        a.item_stride = ((a.rows + (a.row_pad ? 7 : 0)) * a.wpitch + 127) / 128 * 128;     // TMA destinations are 128-byte aligned
        a.raw_w = a.wpitch; a.raw_item_stride = 0; a.aligned16 = 1;
        auto smem_for = [&](int si, int ns) { return (size_t)ns * si * (4 * a.item_stride + bs * bs) + 256 + (size_t)ME_MAX_STAGES * si * 32 + (size_t)((si * tasks_per_item + 31) / 32) * 128; };
        while (SI > 1 && smem_for(SI, a.nstage) > 226 * 1024) --SI;
        smem = smem_for(SI, a.nstage);
    } else {
        a.nstage = 2;
        {   // copies are written by threads (16-B alignment is enough).  G-1 spare rows let the search read past the
            // window without a guard (only invalid candidates see them); the stride is padded so that consecutive tasks
            // keep hitting consecutive 16-byte bank groups across item boundaries: item_stride/16 == NG*G*wpitch/16 (mod 8)
            int is16 = (a.rows + G - 1) * wp16;
            const int want = (a.NG * G * wp16) % 8;
            while (is16 % 8 != want) ++is16;
            a.item_stride = is16 * 16;
        }
        {   // raw TMA box: aligned -> 64-byte rows (3 + 4*NW bytes needed); otherwise up to 15 + 3 more bytes
            const int need_al = 3 + 4 * NW;
            a.aligned16 = (bs % 16 == 0 && r_real % 16 == 0 && need_al <= 64) ? 1 : 0;
            a.raw_w = a.aligned16 ? 64 : (18 + 4 * NW + 15) / 16 * 16;
        }
        a.raw_item_stride = (a.rows * a.raw_w + 127) / 128 * 128;
        auto smem_for = [&](int si, int ns) {
            // 128-byte alignment of the raw area: copies and cur tiles are multiples of 16 only
            return (size_t)ns * si * (4 * a.item_stride + a.raw_item_stride + bs * bs) + 512 + (size_t)ME_MAX_STAGES * si * 32 +
                   (size_t)((si * tasks_per_item + 31) / 32) * 128;
        };
        while (SI > 1 && smem_for(SI, a.nstage) > 226 * 1024) --SI;
        smem = smem_for(SI, a.nstage);
    }
    a.SI = SI;
    a.shift_stride = SI * a.item_stride;
    a.stage_bytes = 4 * a.shift_stride;
    a.raw_stage_bytes = SI * a.raw_item_stride;
    a.NB = (SI * tasks_per_item + 31) / 32;
    a.stages_per_unit = (a.items_per_unit + SI - 1) / SI;
    a.z_per_unit = ctx->nslots * 16;
    for (int i = 0; i < SO_MAX_REF; ++i) a.slot[i] = i < (int)ctx->rs().list.size() ? ctx->rs().list[i] : 0;
    // 1 producer warp + up to 11 search warps.  No more search warps than bundles per stage: a warp may then be at most one
    // stage ahead of the slowest one, which is what the 1-bit mbarrier phase parity can distinguish.
    const int threads = 32 * (1 + (a.NB < 11 ? a.NB : 11));
    CUtensorMap map, cmap;
    int rc = make_ring_map(ctx, a.raw_w, a.rows + (a.row_pad ? 7 : 0), &map);
    if (rc) return rc;
    if (a.direct) {     // current blocks: {W, H, units} view of the frames of this launch
        rc = make_map3d(ctx, cur + (size_t)unit0 * cur_stride, ctx->g.W, ctx->g.H, units, ctx->g.W, cur_stride ? cur_stride : ctx->frame_px, bs, bs, &cmap);
        if (rc) return rc;
    } else {
        cmap = map;     // unused
    }
    // the map addresses the whole ring; unit0 offsets the z coordinate through a.units / unit index in the kernel
    if (unit0 != 0) {      // kernels index units from 0: shift the plane index instead of the base pointer
        for (int i = 0; i < SO_MAX_REF; ++i) a.slot[i] += unit0 * ctx->nslots;
    }
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
    const int total_stages = units * a.stages_per_unit;
    const int grid = total_stages < sms ? total_stages : sms;
    // the key arrays are all-ones here: initialised by keys_reset() and restored by the finish kernel after it decodes them
    const bool quad = out_sub && a.direct && bs == 16 && NDX == 9 && G == 3;
    if (used_quad) *used_quad = quad;
    if (quad) {
        a.out_sub = reinterpret_cast<unsigned long long*>(out_sub + (size_t)unit0 * out_sub_stride);
        a.out_sub_unit_stride = out_sub_stride;
    }
    ev_pair(ctx, ctx->ev_me, st, true);
    if (ctx->timing_on) ctx->ev_xs.push_back(ctx->ev_me.size() - 1);
    cudaError_t e;
    if (quad) e = launch_me_tma<16, 9, 3, true>(map, cmap, a, grid, threads, smem, st);
    else if (bs == 16) e = launch_me_tma_n<16>(NDX, G, map, cmap, a, grid, threads, smem, st);
    else if (bs == 8) e = launch_me_tma_n<8>(NDX, G, map, cmap, a, grid, threads, smem, st);
    else e = launch_me_tma_n<4>(NDX, G, map, cmap, a, grid, threads, smem, st);
    ev_pair(ctx, ctx->ev_me, st, false);
    if (e != cudaSuccess) { set_err(ctx, std::string("me_tma_kernel: ") + cudaGetErrorString(e)); return SO_E_CUDA; }
    ctx->launches++;
    CU(cudaGetLastError());
    return SO_OK;
}

static FlowArgs make_flow(so_ctx* ctx, const uint8_t* cur, size_t cur_stride, const so_frame_out* o, size_t out_frames_stride,
                          int unit0, int qp_rd) {
    // out_frames_stride: number of frames between consecutive units in the output arrays
    FlowArgs a{};
    a.g = ctx->g;
    a.g.nref = (int)ctx->rs().list.size();
    a.ring = make_ring(ctx);
    a.cur = cur; a.cur_unit_stride = cur_stride;
    a.unit0 = unit0;
    a.vbs = (ctx->p.flags & SO_FLAG_VBS) ? 1 : 0;
    a.fast = 0; a.chain = 0; a.nref_fast = ctx->p.n_ref_frames;
    a.qp_final = ctx->p.qp;
    a.qp_rd = qp_rd;
    a.qp_rows = (ctx->p.rc_flag > 0) ? ctx->qp_rows_dev : nullptr;
    a.qp_blocks = ctx->cur_qp_blocks;
    a.lam = ctx->p.lam;
    a.me_parent = ctx->me_parent; a.me_sub = ctx->me_sub;
    a.me_parent_stride = ctx->nblk; a.me_sub_stride = (size_t)ctx->nblk * 4;
    a.res_frame = ctx->res_frame; a.band = ctx->band;
    a.split = o->split; a.mv = o->mv; a.levels = o->levels; a.recon = o->recon; a.row_sizes = o->row_sizes; a.stats = o->stats;
    a.split_stride = out_frames_stride * ctx->nblk;
    a.mv_stride = out_frames_stride * ctx->nblk * 12;
    a.frame_stride = out_frames_stride * ctx->frame_px;
    a.rows_stride = out_frames_stride * ctx->g.nby;
    a.stats_stride = out_frames_stride;
    a.scratch_stride = ctx->frame_px;
    a.blk_len = ctx->cur_blk_len;
    a.blk_len_stride = out_frames_stride * ctx->nblk;
    return a;
}

__global__ void stats_init_kernel(so_frame_stats* st, size_t stride, uint32_t* rows, size_t rows_stride, int nrows, uint32_t den, uint32_t type) {
    so_frame_stats* s = st + blockIdx.x * stride;
    if (threadIdx.x == 0) { s->sse = 0; s->mae_num = 0; s->mae_den = den; s->mae_inf = 0; s->qsize = 0; s->frame_type = type; }
    for (int i = threadIdx.x; i < nrows; i += blockDim.x) rows[blockIdx.x * rows_stride + i] = 0;
}

static inline int nthreads_px(int bs) { return bs * bs < 32 ? 32 : bs * bs; }

static int check_out(so_ctx* ctx, const so_frame_out* o) {
    if (!o || !o->split || !o->mv || !o->levels || !o->recon || !o->row_sizes || !o->stats) {
        set_err(ctx, "so_frame_out has null members"); return SO_E_INVALID;
    }
    if (ctx->p.rc_flag > 0 && ctx->qp_rows.empty()) { set_err(ctx, "rc_flag > 0 requires so_set_row_qps"); return SO_E_STATE; }
    return SO_OK;
}

static int encode_intra_impl(so_ctx* ctx, const uint8_t* cur, size_t cur_stride, const so_frame_out* o, size_t ofs,
                             int unit0, int units, int qp_rd, cudaStream_t st) {
    const FrameGeom& g = ctx->g;
    FlowArgs a = make_flow(ctx, cur, cur_stride, o, ofs, unit0, qp_rd);
    a.me_parent = ctx->in_parent; a.me_sub = ctx->in_sub;
    a.mae_den = (uint32_t)(g.bs * g.bs); a.frame_type = 0;
    if (!ctx->stats_prezeroed) {
        stats_init_kernel<<<units, 64, 0, st>>>(o->stats + unit0 * a.stats_stride, a.stats_stride, o->row_sizes + unit0 * a.rows_stride,
                                                a.rows_stride, g.nby, a.mae_den, 0u);
        ctx->launches++;
    }
    dim3 grid(ctx->nblk, units);
    const int nt = nthreads_px(g.bs);
    const size_t ism = sizeof(unsigned) * 4 * (2 * g.r + 1);
    ev_pair(ctx, ctx->ev_me, st, true);
    static const bool intra_generic = std::getenv("SO_INTRA_GENERIC") != nullptr;     // tests: force the generic intra kernels
    const bool fast16 = g.bs == 16 && g.W % 16 == 0 && !intra_generic;
    if (fast16) intra_search16_kernel<<<dim3((ctx->nblk + 3) / 4, units), 128, 0, st>>>(a);
    else if (g.bs == 16) intra_search_kernel<16><<<grid, 256, ism, st>>>(a);
    else if (g.bs == 8) intra_search_kernel<8><<<grid, 128, ism, st>>>(a);
    else intra_search_kernel<4><<<grid, 64, ism, st>>>(a);
    ev_pair(ctx, ctx->ev_me, st, false);
    ev_pair(ctx, ctx->ev_tq, st, true);
    if (g.bs == 16) intra_finish_kernel<16><<<grid, nt, 0, st>>>(a);
    else if (g.bs == 8) intra_finish_kernel<8><<<grid, nt, 0, st>>>(a);
    else intra_finish_kernel<4><<<grid, nt, 0, st>>>(a);
    dim3 grid2(g.nby, units);
    if (fast16) intra_recon16_kernel<<<grid2, 256, (size_t)g.nbx * 9 + 16, st>>>(a);
    else if (g.bs == 16) intra_recon_kernel<16><<<grid2, nt, 0, st>>>(a);
    else if (g.bs == 8) intra_recon_kernel<8><<<grid2, nt, 0, st>>>(a);
    else intra_recon_kernel<4><<<grid2, nt, 0, st>>>(a);
    ev_pair(ctx, ctx->ev_tq, st, false);
    ctx->launches += 3;
    CU(cudaGetLastError());
    return SO_OK;
}

// units: how many units of the selected chain(s) to touch (lock step: units 0 .. units-1; single unit: 1)
static int ensure_planes(so_ctx* ctx, int units, cudaStream_t st) {
    so_ctx::RingState& R = ctx->rs();
    bool all_u8 = true;
    for (int s : R.list) all_u8 = all_u8 && R.slot_u8[s];
    const FrameGeom& g = ctx->g;
    const int wrap = (g.fme && all_u8) ? 1 : 0;      // np.copy(list) is uint8 only if every frame is (quirk Q1)
    for (int s : R.list) {
        if (!R.slot_u8[s]) continue;                 // the constant frame: every plane was filled at reset
        if (R.slot_wrap[s] == wrap) continue;
        uint8_t* base = slot_ptr(ctx, s) + (size_t)ctx->u0() * ctx->unit_stride;
        ring_planes_kernel<<<dim3((g.W / 4 + 127) / 128, g.H, units), 128, 0, st>>>(base, ctx->unit_stride, ctx->plane_bytes,
                                                                                  base, ctx->unit_stride, g.pitch,
                                                                                  g.W, g.H, g.pitch, g.fme, wrap, 0);
        ctx->launches++;
        R.slot_wrap[s] = wrap;
        R.slot_sea[s] = 0;
    }
    CU(cudaGetLastError());
    return SO_OK;
}

static int encode_inter_impl(so_ctx* ctx, const uint8_t* cur, size_t cur_stride, const so_frame_out* o, size_t ofs,
                             int unit0, int units, cudaStream_t st) {
    const FrameGeom& g = ctx->g;
    if (ctx->rs().list.empty()) { set_err(ctx, "inter frame with an empty reference list (call so_ref_reset)"); return SO_E_STATE; }
    int rc = ensure_planes(ctx, ctx->un(), st);
    if (rc) return rc;
    FlowArgs a = make_flow(ctx, cur, cur_stride, o, ofs, unit0, ctx->p.qp);
    const bool parallel = ctx->p.parallel_mode != 0;
    const bool use_fast = (ctx->p.flags & SO_FLAG_FAST_ME) && ctx->p.parallel_mode != 1;      // Encoder.py:641
    a.fast = use_fast ? 1 : 0;
    a.me_packed = use_fast ? 0 : 1;
    a.chain = parallel ? 0 : 1;
    a.nref_fast = parallel ? 1 : ctx->p.n_ref_frames;                                        // Encoder.py:590
    a.mae_den = use_fast ? 4u : (uint32_t)(g.bs * g.bs); a.frame_type = 1;
    if (!ctx->stats_prezeroed) {
        stats_init_kernel<<<units, 64, 0, st>>>(o->stats + unit0 * a.stats_stride, a.stats_stride, o->row_sizes + unit0 * a.rows_stride,
                                                a.rows_stride, g.nby, a.mae_den, 1u);
        ctx->launches++;
    }
    const int nt = nthreads_px(g.bs);
    if (use_fast) {
        ev_pair(ctx, ctx->ev_me, st, true);
        dim3 grid(a.chain ? 1 : ctx->nblk, units);
        static const bool fast_generic = std::getenv("SO_FAST_GENERIC") != nullptr;      // tests: force the generic kernel
        static const bool fast_no_table = std::getenv("SO_FAST_NO_TABLE") != nullptr;    // A/B: chained fast_me16_kernel
        const bool packed = (g.bs == 16 || g.bs == 8) && g.W % g.bs == 0 && !fast_generic;      // word-packed kernels: 16x16 and 8x8
        if (packed && a.chain && !fast_no_table) {
            // table-driven chain: transition tables around last frame's predictors -> one-warp walk -> parallel results
            const size_t nblk_pad = ((size_t)ctx->nblk + 3) & ~(size_t)3;           // the chain stages groups of four blocks
            if (!ctx->fm_table) {
                CU(cudaMalloc(&ctx->fm_table, (size_t)ctx->batch * nblk_pad * FT_TRANS));
                CU(cudaMalloc(&ctx->fm_state, (size_t)ctx->batch * nblk_pad * sizeof(short4)));
                CU(cudaMemsetAsync(ctx->fm_table, 0xFF, (size_t)ctx->batch * nblk_pad * FT_TRANS, st));
                CU(cudaMemsetAsync(ctx->fm_state, 0, (size_t)ctx->batch * nblk_pad * sizeof(short4), st));
            }
            const int nr = std::min(a.nref_fast, a.g.nref), nph = a.g.fme ? 4 : 1;
            const size_t tsm = (size_t)nr * nph * FTR_H * FTR_W + 256 + (size_t)nr * FT_N * FT_N * 2;
            FlowArgs b = a;
            b.chain = 0; b.mvp_in = ctx->fm_state; b.mvp_in_stride = nblk_pad;
            // the chain itself: a scan over composed transition tables (chunks of fs_L blocks, <= 256 chunks), or -- A/B switch and
            // cross-check in the tests -- the one-warp walk over all blocks
            static const bool chain_walk = std::getenv("SO_FAST_CHAIN_WALK") != nullptr;
            if (!ctx->fs_F) {
                ctx->fs_L = std::max(32, (ctx->nblk + 255) / 256);
                ctx->fs_nchunks = (ctx->nblk + ctx->fs_L - 1) / ctx->fs_L;
                CU(cudaMalloc(&ctx->fs_F, (size_t)ctx->batch * ctx->fs_nchunks * FS_ROW * sizeof(int2)));
                CU(cudaMalloc(&ctx->fs_entry, (size_t)ctx->batch * ctx->fs_nchunks * sizeof(int2)));
                CU(cudaFuncSetAttribute(fast_scan_walk_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                CU(cudaFuncSetAttribute(fast_scan_walk_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            }
            const int L = ctx->fs_L, nch = ctx->fs_nchunks;
            const size_t tstride = nblk_pad * FT_TRANS, fstride = (size_t)nch * FS_ROW;
            const size_t csm = (size_t)L * FT_TRANS + (size_t)L * 4, wsm = (size_t)nch * FS_ROW * sizeof(int2) + (size_t)L * 4;
            // without VBS the whole-block results are all that is needed and the scan knows them: no search kernel afterwards
            static const bool always_me16 = std::getenv("SO_FAST_ALWAYS_ME16") != nullptr;      // A/B switch and cross-check in the tests
            const bool results_from_scan = !chain_walk && !a.vbs && !always_me16;
            auto scan = [&](auto walk_kernel) {
                fast_scan_chunk_kernel<<<dim3(nch, units), 96, csm, st>>>(ctx->fm_table, tstride, ctx->fm_state, nblk_pad, a.unit0, ctx->nblk, L, ctx->fs_F, fstride);
                walk_kernel<<<dim3(1, units), 576, wsm, st>>>(a, ctx->fm_table, tstride, ctx->fm_state, nblk_pad, L, nch, ctx->fs_F, fstride, ctx->fs_entry, (size_t)nch,
                                                              results_from_scan ? 1 : 0);
                fast_scan_fill_kernel<<<dim3(nch, units), 32, csm, st>>>(ctx->fm_table, tstride, ctx->fm_state, nblk_pad, a.unit0, ctx->nblk, L, ctx->fs_entry, (size_t)nch,
                                                                        results_from_scan ? ctx->me_parent : nullptr, (size_t)ctx->nblk);
                ctx->launches += 2;
            };
            if (g.bs == 16) {
                fast_table16_kernel<16><<<dim3(ctx->nblk, units), 128, tsm, st>>>(a, ctx->fm_table, tstride, ctx->fm_state, nblk_pad);
                if (chain_walk) fast_chain16_kernel<16><<<dim3(1, units), 576, 0, st>>>(a, ctx->fm_table, tstride, ctx->fm_state, nblk_pad);
                else scan(fast_scan_walk_kernel<16>);
                if (!results_from_scan) fast_me16_kernel<16><<<dim3(ctx->nblk, units), 576, 0, st>>>(b);
            } else {
                fast_table16_kernel<8><<<dim3(ctx->nblk, units), 128, tsm, st>>>(a, ctx->fm_table, tstride, ctx->fm_state, nblk_pad);
                if (chain_walk) fast_chain16_kernel<8><<<dim3(1, units), 576, 0, st>>>(a, ctx->fm_table, tstride, ctx->fm_state, nblk_pad);
                else scan(fast_scan_walk_kernel<8>);
                if (!results_from_scan) fast_me16_kernel<8><<<dim3(ctx->nblk, units), 576, 0, st>>>(b);
            }
            if (results_from_scan) ctx->launches--;
            ctx->launches += 2;
        }
        else if (packed && g.bs == 16) fast_me16_kernel<16><<<grid, 576, 0, st>>>(a);
        else if (packed) fast_me16_kernel<8><<<grid, 576, 0, st>>>(a);
        else if (g.bs == 16) fast_me_kernel<16><<<grid, nt, 0, st>>>(a);
        else if (g.bs == 8) fast_me_kernel<8><<<grid, nt, 0, st>>>(a);
        else fast_me_kernel<4><<<grid, nt, 0, st>>>(a);
        ev_pair(ctx, ctx->ev_me, st, false);
        ctx->launches++;
    } else {
        bool quad = false;
        rc = run_me_tma(ctx, cur, cur_stride, unit0, units, g.bs, ctx->me_parent, ctx->nblk, st,
                        a.vbs ? ctx->me_sub : nullptr, (size_t)ctx->nblk * 4, &quad);
        if (rc) return rc;
        if (a.vbs && !quad) {       // generic geometry: second search on the sub-block grid
            rc = run_me_tma(ctx, cur, cur_stride, unit0, units, g.bs / 2, ctx->me_sub, (size_t)ctx->nblk * 4, st);
            if (rc) return rc;
        }
    }
    if (!use_fast) a.me_packed = ctx->me_key_fmt;
    ev_pair(ctx, ctx->ev_tq, st, true);
    if (ctx->timing_on) ctx->ev_tq_inter.push_back(ctx->ev_tq.size() - 1);
    dim3 grid(ctx->nblk, units);
    static const bool generic16 = std::getenv("SO_FINISH_GENERIC") != nullptr;      // tests: force the generic kernel
    if (g.bs == 16 && g.W % 16 == 0 && !generic16) {
        const dim3 fg((ctx->nblk + 7) / 8, units);
        if (a.vbs) CU(launch_pdl(inter_finish16_kernel<true>, fg, dim3(128), 0, st, a));
        else CU(launch_pdl(inter_finish16_kernel<false>, fg, dim3(128), 0, st, a));
    }
    else if (g.bs == 16) inter_finish_kernel<16><<<grid, nt, 0, st>>>(a);
    else if (g.bs == 8) inter_finish_kernel<8><<<grid, nt, 0, st>>>(a);
    else inter_finish_kernel<4><<<grid, nt, 0, st>>>(a);
    ev_pair(ctx, ctx->ev_tq, st, false);
    ctx->launches++;
    CU(cudaGetLastError());
    return SO_OK;
}

// Per-frame seam (complete_intra_flow Encoder.py:1582 / complete_inter_flow :1644).  unit >= 0: ONE frame of that unit's
// chain (cur / out are single-frame buffers); SO_ALL_UNITS: one frame of every unit in lock step (cur and every output
// plane are dense [max_batch][...] arrays).
static int frame_call(so_ctx* ctx, int unit, const uint8_t* cur_dev, const so_frame_out* out, cudaStream_t st, bool intra) {
    if (!ctx || !cur_dev) return SO_E_INVALID;
    int rc = check_out(ctx, out);
    if (rc) return rc;
    CU(cudaSetDevice(ctx->device));
    rc = select_unit(ctx, unit);
    if (rc) return rc;
    ctx->timing_on = false;
    const bool one = unit >= 0;
    // single unit: every per-unit stride is zero, so the kernels' `unit * stride` terms vanish and the buffers are the frame
    const size_t cur_stride = one ? 0 : ctx->frame_px, ofs = one ? 0 : 1;
    if (intra) return encode_intra_impl(ctx, cur_dev, cur_stride, out, ofs, ctx->u0(), ctx->un(), ctx->p.qp, st);
    return encode_inter_impl(ctx, cur_dev, cur_stride, out, ofs, ctx->u0(), ctx->un(), st);
}

extern "C" int so_encode_intra(so_ctx* ctx, int unit, const uint8_t* cur_dev, const so_frame_out* out, void* stream) {
    return frame_call(ctx, unit, cur_dev, out, (cudaStream_t)stream, true);
}

extern "C" int so_encode_inter(so_ctx* ctx, int unit, const uint8_t* cur_dev, const so_frame_out* out, void* stream) {
    return frame_call(ctx, unit, cur_dev, out, (cudaStream_t)stream, false);
}

// ---------------------------------------------------------------------------------------------------------
// sequence encode (frame loop of Encoder.py:1829-1871)
// ---------------------------------------------------------------------------------------------------------
static int ensure_seq(so_ctx* ctx, size_t nframes_total) {
    if (nframes_total <= ctx->sq_cap_frames) return SO_OK;
    free_seq(ctx);
    const size_t n = nframes_total;
    CU(cudaMalloc(&ctx->sq_frames, n * ctx->frame_px));
    CU(cudaMalloc(&ctx->sq_recon, n * ctx->frame_px));
    CU(cudaMalloc(&ctx->sq_levels, n * ctx->frame_px * sizeof(int16_t)));
    CU(cudaMalloc(&ctx->sq_split, n * ctx->nblk));
    CU(cudaMalloc(&ctx->sq_mv, n * ctx->nblk * 12 * sizeof(int16_t)));
    CU(cudaMalloc(&ctx->sq_rows, n * ctx->g.nby * sizeof(uint32_t)));
    CU(cudaMalloc(&ctx->sq_blklen, n * ctx->nblk * sizeof(uint32_t)));
    CU(cudaMalloc(&ctx->sq_stats, n * sizeof(so_frame_stats)));
    ctx->sq_cap_frames = n;
    return SO_OK;
}

// ---- chunk pipeline helpers ------------------------------------------------------------------------------------
static int ensure_sym(so_ctx* ctx, size_t total);
static int sym_run_range(so_ctx* ctx, int f0, int nf);
static int pipe_symbols_chunk(so_ctx* ctx, int c);

static int pipe_upload_chunk(so_ctx* ctx, int c) {
    auto& P = ctx->pipe;
    const int U = ctx->sq_units, F = ctx->sq_nframes;
    const int f0 = c * P.chunk, n = std::min(P.chunk, F - f0);
    const size_t px = ctx->frame_px;
    if (P.fd >= 0) {
        // Y extraction + padding (read_yuv Encoder.py:110-126, pad_hw :140-155): the chroma planes are skipped on disk, rows and
        // columns beyond the source size are 128.  The host blocks in pread() here while the GPU works through the two chunks
        // that are already queued.
        uint8_t* dst = P.stage[c % 3];
        if (c >= 3) CU(cudaEventSynchronize(P.up[c - 3]));            // the H2D copy that last read this staging buffer
        const int W = ctx->g.W, H = ctx->g.H;
        const size_t ysz = (size_t)P.src_w * P.src_h, fsz = ysz + 2 * (ysz / 4);
        const bool padded = P.src_w != W || P.src_h != H;
        if (padded) memset(dst, 128, (size_t)n * px);
        for (int i = 0; i < n; ++i) {
            const off_t base = (off_t)((size_t)(P.first_frame + f0 + i) * fsz);
            if (!padded) {
                size_t got = 0;
                while (got < ysz) {
                    const ssize_t r = pread(P.fd, dst + (size_t)i * px + got, ysz - got, base + (off_t)got);
                    if (r <= 0) { set_err(ctx, "yuv file is too short for the requested frames"); return SO_E_INVALID; }
                    got += (size_t)r;
                }
            } else {
                for (int y = 0; y < P.src_h; ++y) {
                    size_t got = 0;
                    while (got < (size_t)P.src_w) {
                        const ssize_t r = pread(P.fd, dst + (size_t)i * px + (size_t)y * W + got, (size_t)P.src_w - got,
                                                base + (off_t)((size_t)y * P.src_w + got));
                        if (r <= 0) { set_err(ctx, "yuv file is too short for the requested frames"); return SO_E_INVALID; }
                        got += (size_t)r;
                    }
                }
            }
        }
        CU(cudaMemcpyAsync(ctx->sq_frames + (size_t)f0 * px, dst, (size_t)n * px, cudaMemcpyHostToDevice, ctx->st_h2d));
        CU(cudaEventRecord(P.up[c], ctx->st_h2d));
        return SO_OK;
    }
    for (int u = 0; u < U; ++u)
        CU(cudaMemcpyAsync(ctx->sq_frames + ((size_t)u * F + f0) * px, P.h_frames + ((size_t)u * F + f0) * px, (size_t)n * px,
                           cudaMemcpyHostToDevice, ctx->st_h2d));
    CU(cudaEventRecord(P.up[c], ctx->st_h2d));
    return SO_OK;
}

static int pipe_download_chunk(so_ctx* ctx, int c) {
    auto& P = ctx->pipe;
    const int U = ctx->sq_units, F = ctx->sq_nframes;
    const int f0 = c * P.chunk, n = std::min(P.chunk, F - f0);
    const size_t px = ctx->frame_px, nblk = ctx->nblk, nby = ctx->g.nby;
    cudaStream_t s = ctx->st_d2h;
    CU(cudaStreamWaitEvent(s, P.done[c], 0));
    for (int u = 0; u < U; ++u) {
        const size_t o = (size_t)u * F + f0;
        CU(cudaMemcpyAsync(P.h_split + o * nblk, ctx->sq_split + o * nblk, n * nblk, cudaMemcpyDeviceToHost, s));
        CU(cudaMemcpyAsync(P.h_mv + o * nblk * 12, ctx->sq_mv + o * nblk * 12, n * nblk * 12 * sizeof(int16_t), cudaMemcpyDeviceToHost, s));
        if (P.h_levels) CU(cudaMemcpyAsync(P.h_levels + o * px, ctx->sq_levels + o * px, n * px * sizeof(int16_t), cudaMemcpyDeviceToHost, s));
        if (P.h_recon) CU(cudaMemcpyAsync(P.h_recon + o * px, ctx->sq_recon + o * px, n * px, cudaMemcpyDeviceToHost, s));
        if (P.h_rows) CU(cudaMemcpyAsync(P.h_rows + o * nby, ctx->sq_rows + o * nby, n * nby * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        CU(cudaMemcpyAsync(P.h_stats + o, ctx->sq_stats + o, n * sizeof(so_frame_stats), cudaMemcpyDeviceToHost, s));
        if (ctx->sym_out) CU(cudaMemcpyAsync(ctx->h_tot + o, ctx->sym_tot + o, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    }
    if (ctx->sym_out) CU(cudaEventRecord(P.tot[c], s));
    return SO_OK;
}

// stage 1: host -> device copy of the input frames (async on the context stream)
extern "C" int so_seq_upload(so_ctx* ctx, const uint8_t* frames, int n_units, int n_frames) {
    if (!ctx || !frames || n_units < 1 || n_frames < 1) return SO_E_INVALID;
    if (n_units > ctx->batch) { set_err(ctx, "n_units exceeds max_batch of the context"); return SO_E_INVALID; }
    CU(cudaSetDevice(ctx->device));
    const size_t total = (size_t)n_units * n_frames;
    int rc = ensure_seq(ctx, total);
    if (rc) return rc;
    CU(cudaMemcpyAsync(ctx->sq_frames, frames, total * ctx->frame_px, cudaMemcpyHostToDevice, ctx->stream));
    ctx->sq_units = n_units; ctx->sq_nframes = n_frames;
    return SO_OK;
}

// stage 2: the frame loop on frames already resident in HBM; outputs stay on the device.  Asynchronous unless
// rc_flag == 2 (the scene-cut decision needs quantized_sized on the host every P frame).
extern "C" int so_seq_run(so_ctx* ctx) {
    if (!ctx) return SO_E_INVALID;
    if (ctx->sq_units < 1) { set_err(ctx, "so_seq_run before so_seq_upload"); return SO_E_STATE; }
    if (ctx->p.rc_flag > 0 && ctx->qp_rows.empty()) { set_err(ctx, "rc_flag > 0 requires so_set_row_qps"); return SO_E_STATE; }
    CU(cudaSetDevice(ctx->device));
    const int n_units = ctx->sq_units, n_frames = ctx->sq_nframes;
    cudaStream_t st = ctx->stream;
    const size_t px = ctx->frame_px;
    const int nby = ctx->g.nby;
    ctx->launches = 0;
    ctx->ev_used = 0; ctx->ev_me.clear(); ctx->ev_tq.clear(); ctx->ev_xs.clear(); ctx->ev_tq_inter.clear();
    ctx->timing_on = true;
    while (ctx->ev_pool.size() < 2) { cudaEvent_t e; CU(cudaEventCreate(&e)); ctx->ev_pool.push_back(e); }
    ctx->ev0 = ctx->ev_pool[ctx->ev_used++]; ctx->ev1 = ctx->ev_pool[ctx->ev_used++];
    CU(cudaEventRecord(ctx->ev0, st));
    int rc = so_ref_reset(ctx, SO_ALL_UNITS, st);
    if (rc) return rc;
    CU(cudaMemsetAsync(ctx->sq_stats, 0, (size_t)n_units * n_frames * sizeof(so_frame_stats), st));
    CU(cudaMemsetAsync(ctx->sq_rows, 0, (size_t)n_units * n_frames * nby * sizeof(uint32_t), st));
    ctx->stats_prezeroed = true;
    struct Unflag { so_ctx* c; ~Unflag() { c->stats_prezeroed = false; } } unflag{ctx};
    std::vector<so_frame_stats> hstats(n_units);
    // per-kernel CUDA events cost ~10 us of stream serialisation per frame (1.8 % of a 1080p P frame), so they are recorded on
    // every timing_stride-th frame only (default 8; SO_TIMING_STRIDE=1 times every launch)
    static const int timing_stride = std::getenv("SO_TIMING_STRIDE") ? std::max(1, atoi(std::getenv("SO_TIMING_STRIDE"))) : 8;
    ctx->timed_frames = 0; ctx->total_frames = n_frames;
    for (int f = 0; f < n_frames; ++f) {
        ctx->timing_on = (f % timing_stride) == 0;
        if (ctx->timing_on) ctx->timed_frames++;
        if (ctx->pipe.active && f % ctx->pipe.chunk == 0) {
            const int c = f / ctx->pipe.chunk;
            if (c + 2 < ctx->pipe.nchunks) { rc = pipe_upload_chunk(ctx, c + 2); if (rc) return rc; }   // two chunks ahead
            CU(cudaStreamWaitEvent(st, ctx->pipe.up[c], 0));
        }
        so_frame_out o;
        o.split = ctx->sq_split + (size_t)f * ctx->nblk;
        o.mv = ctx->sq_mv + (size_t)f * ctx->nblk * 12;
        o.levels = ctx->sq_levels + (size_t)f * px;
        o.recon = ctx->sq_recon + (size_t)f * px;
        o.row_sizes = ctx->sq_rows + (size_t)f * nby;
        o.stats = ctx->sq_stats + f;
        const uint8_t* cur = ctx->sq_frames + (size_t)f * px;
        const size_t cur_stride = (size_t)n_frames * px;
        const bool intra = (f % ctx->p.intra_dur == 0) && ctx->p.parallel_mode != 1;       // Encoder.py:1839
        ctx->cur_qp_blocks = (ctx->qp_blocks_dev && f < ctx->qp_blocks_frames) ? ctx->qp_blocks_dev + (size_t)f * ctx->nblk : nullptr;
        ctx->cur_blk_len = ctx->sq_blklen + (size_t)f * ctx->nblk;
        struct ClearQ { so_ctx* c; ~ClearQ() { c->cur_qp_blocks = nullptr; c->cur_blk_len = nullptr; } } clearq{ctx};
        if (intra) {
            rc = encode_intra_impl(ctx, cur, cur_stride, &o, n_frames, 0, n_units, ctx->p.qp, st);
            if (rc) return rc;
        } else {
            if (ctx->p.parallel_mode == 1) { rc = ref_reset_impl(ctx, st, false); if (rc) return rc; }   // Encoder.py:1846
            rc = encode_inter_impl(ctx, cur, cur_stride, &o, n_frames, 0, n_units, st);
            if (rc) return rc;
            if (ctx->p.rc_flag > 1) {                                                     // scene cut, Encoder.py:1851-1856
                CU(cudaStreamSynchronize(st));
                for (int u = 0; u < n_units; ++u)
                    CU(cudaMemcpy(&hstats[u], o.stats + (size_t)u * n_frames, sizeof(so_frame_stats), cudaMemcpyDeviceToHost));
                for (int u = 0; u < n_units; ++u) {
                    if ((int64_t)hstats[u].qsize > ctx->p.intra_thresh) {
                        // self.Qp still holds the last row QP of the inter flow: it is the RD QP of this intra pass
                        ctx->stats_prezeroed = false;        // the frame's statistics hold the inter attempt: re-initialise
                        rc = encode_intra_impl(ctx, cur, cur_stride, &o, n_frames, u, 1, ctx->qp_rows.back(), st);
                        ctx->stats_prezeroed = true;
                        if (rc) return rc;
                    }
                }
            }
        }
        if (f < n_frames - 1) {
            rc = ring_push(ctx, o.recon, (size_t)n_frames * px, n_units, st);
            if (rc) return rc;
        }
        if (ctx->pipe.active && ((f + 1) % ctx->pipe.chunk == 0 || f == n_frames - 1)) {
            const int c = f / ctx->pipe.chunk;
            if (ctx->sym_out) {     // run-level symbols of the chunk: count -> scan -> emit, then they travel instead of the levels
                rc = sym_run_range(ctx, c * ctx->pipe.chunk, f + 1 - c * ctx->pipe.chunk);
                if (rc) return rc;
            }
            CU(cudaEventRecord(ctx->pipe.done[c], st));
            rc = pipe_download_chunk(ctx, c);
            if (rc) return rc;
            // the symbol copies of the PREVIOUS chunk need its per-frame counts on the host: by now the GPU has a whole chunk
            // of work queued behind them, so this wait throttles the host without starving the device
            if (ctx->sym_out && c > 0) { rc = pipe_symbols_chunk(ctx, c - 1); if (rc) return rc; }
            if (ctx->sym_out && f == n_frames - 1) { rc = pipe_symbols_chunk(ctx, c); if (rc) return rc; }
        }
    }
    CU(cudaEventRecord(ctx->ev1, st));
    ctx->timing_on = false;
    ctx->timing_pending = true;
    return SO_OK;
}

// stage 3: device -> host copy of the outputs (any pointer except split/mv/stats may be NULL) and synchronise
extern "C" int so_seq_download(so_ctx* ctx, uint8_t* split, int16_t* mv, int16_t* levels, uint8_t* recon,
                               uint32_t* row_sizes, so_frame_stats* stats) {
    if (!ctx) return SO_E_INVALID;
    if (ctx->sq_units < 1) { set_err(ctx, "so_seq_download before so_seq_upload"); return SO_E_STATE; }
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const size_t total = (size_t)ctx->sq_units * ctx->sq_nframes, px = ctx->frame_px;
    if (split) CU(cudaMemcpyAsync(split, ctx->sq_split, total * ctx->nblk, cudaMemcpyDeviceToHost, st));
    if (mv) CU(cudaMemcpyAsync(mv, ctx->sq_mv, total * ctx->nblk * 12 * sizeof(int16_t), cudaMemcpyDeviceToHost, st));
    if (levels) CU(cudaMemcpyAsync(levels, ctx->sq_levels, total * px * sizeof(int16_t), cudaMemcpyDeviceToHost, st));
    if (recon) CU(cudaMemcpyAsync(recon, ctx->sq_recon, total * px, cudaMemcpyDeviceToHost, st));
    if (row_sizes) CU(cudaMemcpyAsync(row_sizes, ctx->sq_rows, total * ctx->g.nby * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    if (stats) CU(cudaMemcpyAsync(stats, ctx->sq_stats, total * sizeof(so_frame_stats), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return SO_OK;
}

extern "C" int so_seq_sync(so_ctx* ctx) {
    if (!ctx) return SO_E_INVALID;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    return SO_OK;
}

// Frames per pipeline chunk (the unit of upload, symbol packing and download): 8, fewer when many units are batched so that a
// chunk stays around 256 MB of input -- the first chunk's upload is the only copy the encode cannot hide, and with 64 4K
// units an 8-frame chunk would be 4 GB of it.  SO_PIPE_CHUNK=n overrides (tests, A/B).
static int pipe_chunk_frames(so_ctx* ctx, int n_units) {
    static const int forced = std::getenv("SO_PIPE_CHUNK") ? atoi(std::getenv("SO_PIPE_CHUNK")) : 0;
    if (forced > 0) return forced;
    const size_t per_frame = (size_t)n_units * ctx->frame_px;
    size_t c = ((size_t)256 << 20) / std::max<size_t>(per_frame, 1);
    return (int)std::min<size_t>(8, std::max<size_t>(1, c));
}

// The chunked pipeline shared by so_encode_sequence and so_encode_yuv420_file (pipe.* and sq_units / sq_nframes are set):
// uploads two chunks ahead, the frame loop, downloads one chunk behind; with so_set_symbol_output the packed run-level
// symbols are produced per chunk on the device and copied out with their exact sizes.
static int pipe_run(so_ctx* ctx) {
    auto& P = ctx->pipe;
    while ((int)P.up.size() < P.nchunks) {
        cudaEvent_t a, b, c;
        CU(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&c, cudaEventDisableTiming));
        P.up.push_back(a); P.done.push_back(b); P.tot.push_back(c);
    }
    int rc = SO_OK;
    ctx->sym_resident = false;
    if (ctx->sym_out) {
        rc = ensure_sym(ctx, (size_t)ctx->sq_units * ctx->sq_nframes);
        if (rc) return rc;
        ctx->sym_used = 0; ctx->sym_overflow = false;
        ctx->sym_out->needed = 0;
    }
    P.active = true;
    rc = pipe_upload_chunk(ctx, 0);
    if (!rc && P.nchunks > 1) rc = pipe_upload_chunk(ctx, 1);
    if (!rc) rc = so_seq_run(ctx);
    P.active = false;
    if (rc) { cudaStreamSynchronize(ctx->stream); cudaStreamSynchronize(ctx->st_h2d); cudaStreamSynchronize(ctx->st_d2h); return rc; }
    CU(cudaStreamSynchronize(ctx->st_d2h));
    CU(cudaStreamSynchronize(ctx->stream));
    if (ctx->sym_out) {
        ctx->sym_resident = true;
        ctx->sym_out->needed = ctx->sym_used;
        if (ctx->sym_overflow) {
            set_err(ctx, "symbol buffer too small: every other output is complete; enlarge it to `needed` and call so_fetch_symbols");
            return SO_E_NOMEM;
        }
    }
    return SO_OK;
}

extern "C" int so_encode_sequence(so_ctx* ctx, const uint8_t* frames, int n_units, int n_frames, uint8_t* split, int16_t* mv,
                                  int16_t* levels, uint8_t* recon, uint32_t* row_sizes, so_frame_stats* stats) {
    if (!ctx || !frames || !split || !mv || !stats || n_units < 1 || n_frames < 1) return SO_E_INVALID;
    if (n_units > ctx->batch) { set_err(ctx, "n_units exceeds max_batch of the context"); return SO_E_INVALID; }
    CU(cudaSetDevice(ctx->device));
    int rc = ensure_seq(ctx, (size_t)n_units * n_frames);
    if (rc) return rc;
    ctx->sq_units = n_units; ctx->sq_nframes = n_frames;
    auto& P = ctx->pipe;
    P.chunk = pipe_chunk_frames(ctx, n_units);
    P.nchunks = (n_frames + P.chunk - 1) / P.chunk;
    P.h_frames = frames; P.h_split = split; P.h_mv = mv; P.h_levels = levels; P.h_recon = recon; P.h_rows = row_sizes; P.h_stats = stats;
    return pipe_run(ctx);
}

// Frame ingest from a planar YUV 4:2:0 file (Encoder.py:110-126 read_yuv + :140-155 pad_hw) fused with the encode of one
// sequence: only the luma planes are read, chunk by chunk, into rotating pinned buffers; file reads, H2D copies, the
// encode and the D2H copies of different chunks overlap.
extern "C" int so_encode_yuv420_file(so_ctx* ctx, const char* path, int src_width, int src_height, int first_frame, int n_frames,
                                     uint8_t* split, int16_t* mv, int16_t* levels, uint8_t* recon, uint32_t* row_sizes,
                                     so_frame_stats* stats) {
    if (!ctx || !path || !split || !mv || !stats || n_frames < 1 || first_frame < 0) return SO_E_INVALID;
    if (src_width < 1 || src_height < 1 || src_width > ctx->g.W || src_height > ctx->g.H) {
        set_err(ctx, "source size must be positive and not larger than the coded size of the context");
        return SO_E_INVALID;
    }
    CU(cudaSetDevice(ctx->device));
    int rc = ensure_seq(ctx, (size_t)n_frames);
    if (rc) return rc;
    auto& P = ctx->pipe;
    P.fd = open(path, O_RDONLY);
    if (P.fd < 0) { set_err(ctx, std::string("cannot open ") + path); return SO_E_INVALID; }
    struct Closer { so_ctx::Pipe& p; ~Closer() { if (p.fd >= 0) close(p.fd); p.fd = -1; p.active = false; } } closer{P};
    ctx->sq_units = 1; ctx->sq_nframes = n_frames;
    P.chunk = 8;
    P.nchunks = (n_frames + P.chunk - 1) / P.chunk;
    P.src_w = src_width; P.src_h = src_height; P.first_frame = first_frame;
    const size_t need = (size_t)P.chunk * ctx->frame_px;
    if (P.stage_bytes < need) {
        for (auto& b : P.stage) { if (b) cudaFreeHost(b); b = nullptr; }
        for (auto& b : P.stage) CU(cudaHostAlloc(&b, need, cudaHostAllocDefault));
        P.stage_bytes = need;
    }
    P.h_frames = nullptr; P.h_split = split; P.h_mv = mv; P.h_levels = levels; P.h_recon = recon; P.h_rows = row_sizes; P.h_stats = stats;
    return pipe_run(ctx);
}

// ---------------------------------------------------------------------------------------------------------
// run-level symbols of the resident sequence: count -> device prefix scan -> emit (packed per frame)
// ---------------------------------------------------------------------------------------------------------
static int ensure_sym(so_ctx* ctx, size_t total) {
    if (total <= ctx->sym_cap_frames) return SO_OK;
    const int nblk = ctx->nblk;
    cudaFree(ctx->sym_boffs); cudaFree(ctx->sym_data); cudaFree(ctx->sym_tot);
    if (ctx->h_tot) cudaFreeHost(ctx->h_tot);
    ctx->sym_boffs = ctx->sym_tot = nullptr; ctx->sym_data = nullptr; ctx->h_tot = nullptr; ctx->sym_cap_frames = 0;
    ctx->sym_frame_stride = ctx->frame_px + ctx->frame_px / 2 + (size_t)4 * nblk;      // worst case: 1.5 symbols per coefficient + 1 per (sub-)block
    CU(cudaMalloc(&ctx->sym_boffs, total * (nblk + 1) * sizeof(uint32_t)));
    CU(cudaMalloc(&ctx->sym_tot, total * sizeof(uint32_t)));
    CU(cudaMalloc(&ctx->sym_data, total * ctx->sym_frame_stride * sizeof(int16_t)));
    CU(cudaHostAlloc(&ctx->h_tot, total * sizeof(uint32_t), cudaHostAllocDefault));
    ctx->sym_cap_frames = total;
    return SO_OK;
}

// Symbol packer of the end-to-end path for frames [f0, f0 + nf) of every unit of the resident sequence, on the context
// stream: the per-block counts were written by the finish kernels (sq_blklen); exclusive scan per frame, then one warp per
// block writes its symbols at their final place of the frame's packed stream.
static int sym_run_range(so_ctx* ctx, int f0, int nf) {
    const FrameGeom& g = ctx->g;
    const int nblk = ctx->nblk, F = ctx->sq_nframes, U = ctx->sq_units;
    cudaStream_t st = ctx->stream;
    scan_lens_kernel<<<(unsigned)(U * nf), 1024, 0, st>>>(ctx->sq_blklen, ctx->sym_boffs, nblk, ctx->sym_tot, f0, nf, F);
    dim3 grid((nblk + 7) / 8, (unsigned)(U * nf));
    if (g.bs == 16) rle_emit_warp_kernel<16><<<grid, 256, 0, st>>>(ctx->sq_levels, ctx->sq_split, ctx->sym_boffs, ctx->sym_data, ctx->sym_frame_stride, g.W, g.H, g.nbx, nblk, f0, nf, F);
    else if (g.bs == 8) rle_emit_warp_kernel<8><<<grid, 256, 0, st>>>(ctx->sq_levels, ctx->sq_split, ctx->sym_boffs, ctx->sym_data, ctx->sym_frame_stride, g.W, g.H, g.nbx, nblk, f0, nf, F);
    else rle_emit_warp_kernel<4><<<grid, 256, 0, st>>>(ctx->sq_levels, ctx->sq_split, ctx->sym_boffs, ctx->sym_data, ctx->sym_frame_stride, g.W, g.H, g.nbx, nblk, f0, nf, F);
    ctx->launches += 2;
    CU(cudaGetLastError());
    return SO_OK;
}

// The symbol API with per-sub-block offsets (so_seq_symbols / so_seq_download_symbols): its own count pass (one CTA per
// block), scan and emit over the whole resident sequence.  It shares sym_data / sym_tot with the packer above -- both
// produce the same stream.
extern "C" int so_seq_symbols(so_ctx* ctx) {
    if (!ctx) return SO_E_INVALID;
    if (ctx->sq_units < 1) { set_err(ctx, "so_seq_symbols before a sequence was encoded"); return SO_E_STATE; }
    CU(cudaSetDevice(ctx->device));
    const size_t total = (size_t)ctx->sq_units * ctx->sq_nframes;
    int rc = ensure_sym(ctx, total);
    if (rc) return rc;
    const FrameGeom& g = ctx->g;
    const int nblk = ctx->nblk, F = ctx->sq_nframes;
    if (total > ctx->sym_sub_cap_frames) {
        cudaFree(ctx->sym_lens); cudaFree(ctx->sym_offs);
        ctx->sym_lens = ctx->sym_offs = nullptr; ctx->sym_sub_cap_frames = 0;
        CU(cudaMalloc(&ctx->sym_lens, total * nblk * 4 * sizeof(uint32_t)));
        CU(cudaMalloc(&ctx->sym_offs, total * (nblk * 4 + 1) * sizeof(uint32_t)));
        ctx->sym_sub_cap_frames = total;
    }
    cudaStream_t st = ctx->stream;
    dim3 grid(nblk, (unsigned)total);
    const int nt = nthreads_px(g.bs);
    for (int emit = 0; emit < 2; ++emit) {
        if (g.bs == 16) rle_symbols_kernel<16><<<grid, nt, 0, st>>>(ctx->sq_levels, ctx->sq_split, ctx->sym_lens, ctx->sym_offs, ctx->sym_data, ctx->sym_frame_stride, g.W, g.nbx, nblk, emit, 0, F, F);
        else if (g.bs == 8) rle_symbols_kernel<8><<<grid, nt, 0, st>>>(ctx->sq_levels, ctx->sq_split, ctx->sym_lens, ctx->sym_offs, ctx->sym_data, ctx->sym_frame_stride, g.W, g.nbx, nblk, emit, 0, F, F);
        else rle_symbols_kernel<4><<<grid, nt, 0, st>>>(ctx->sq_levels, ctx->sq_split, ctx->sym_lens, ctx->sym_offs, ctx->sym_data, ctx->sym_frame_stride, g.W, g.nbx, nblk, emit, 0, F, F);
        if (!emit) scan_lens_kernel<<<(unsigned)total, 1024, 0, st>>>(ctx->sym_lens, ctx->sym_offs, nblk * 4, ctx->sym_tot, 0, F, F);
    }
    ctx->launches += 3;
    CU(cudaGetLastError());
    ctx->sym_resident = true;
    return SO_OK;
}

// ---- packed symbols as an output of the sequence encodes (so_set_symbol_output) --------------------------------------
extern "C" int so_set_symbol_output(so_ctx* ctx, so_symbol_out* sym) {
    if (!ctx) return SO_E_INVALID;
    if (sym && (!sym->pos || !sym->count || (!sym->symbols && sym->capacity))) { set_err(ctx, "so_symbol_out needs pos, count and a buffer for its capacity"); return SO_E_INVALID; }
    ctx->sym_out = sym;
    return SO_OK;
}

// the totals of chunk c have been copied to h_tot (event pipe.tot[c]): hand every frame of the chunk its place in the
// caller's buffer and enqueue the copies of exactly that many symbols
static int pipe_symbols_chunk(so_ctx* ctx, int c) {
    auto& P = ctx->pipe;
    so_symbol_out* so = ctx->sym_out;
    const int U = ctx->sq_units, F = ctx->sq_nframes;
    const int f0 = c * P.chunk, n = std::min(P.chunk, F - f0);
    CU(cudaEventSynchronize(P.tot[c]));
    for (int u = 0; u < U; ++u)
        for (int i = 0; i < n; ++i) {
            const size_t fr = (size_t)u * F + f0 + i;
            const uint64_t cnt = ctx->h_tot[fr];
            so->count[fr] = (uint32_t)cnt;
            so->pos[fr] = ctx->sym_used;
            if (ctx->sym_used + cnt <= so->capacity) {
                if (cnt) CU(cudaMemcpyAsync(so->symbols + ctx->sym_used, ctx->sym_data + fr * ctx->sym_frame_stride, cnt * sizeof(int16_t),
                                            cudaMemcpyDeviceToHost, ctx->st_d2h));
            } else {
                ctx->sym_overflow = true;         // keep counting: `needed` tells the caller how large the buffer has to be
            }
            ctx->sym_used += cnt;
        }
    return SO_OK;
}

// After a sequence encode that reported SO_E_NOMEM for the symbols (or simply again): copy the symbols of the resident
// sequence into sym->symbols in (unit, frame) order; fills pos / count / needed.
extern "C" int so_fetch_symbols(so_ctx* ctx, so_symbol_out* sym) {
    if (!ctx || !sym || !sym->pos || !sym->count) return SO_E_INVALID;
    if (!ctx->sym_resident) { set_err(ctx, "no symbol streams resident (encode with so_set_symbol_output, or call so_seq_symbols)"); return SO_E_STATE; }
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const size_t total = (size_t)ctx->sq_units * ctx->sq_nframes;
    CU(cudaMemcpyAsync(ctx->h_tot, ctx->sym_tot, total * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    uint64_t acc = 0;
    for (size_t f = 0; f < total; ++f) { sym->pos[f] = acc; sym->count[f] = ctx->h_tot[f]; acc += ctx->h_tot[f]; }
    sym->needed = acc;
    if (acc > sym->capacity || (!sym->symbols && acc)) { set_err(ctx, "symbol buffer too small"); return SO_E_NOMEM; }
    for (size_t f = 0; f < total; ++f)
        if (sym->count[f]) CU(cudaMemcpyAsync(sym->symbols + sym->pos[f], ctx->sym_data + f * ctx->sym_frame_stride,
                                              (size_t)sym->count[f] * sizeof(int16_t), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return SO_OK;
}

// offsets u32 [units*frames][4*n_blocks + 1] (exclusive, per frame; the last entry is the frame's symbol count);
// symbols i16: frame f of the packed output starts at sym_base[f] (u64 [units*frames + 1], filled here).  Returns
// SO_E_NOMEM with the required symbol count in *needed when sym_capacity is too small.
extern "C" int so_seq_download_symbols(so_ctx* ctx, uint32_t* offsets, int16_t* symbols, uint64_t sym_capacity, uint64_t* sym_base,
                                       uint64_t* needed) {
    if (!ctx || !offsets || !sym_base) return SO_E_INVALID;
    if (!ctx->sym_offs) { set_err(ctx, "so_seq_download_symbols before so_seq_symbols"); return SO_E_STATE; }
    CU(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const size_t total = (size_t)ctx->sq_units * ctx->sq_nframes;
    const size_t n1 = (size_t)ctx->nblk * 4 + 1;
    CU(cudaMemcpyAsync(offsets, ctx->sym_offs, total * n1 * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    uint64_t acc = 0;
    for (size_t f = 0; f < total; ++f) { sym_base[f] = acc; acc += offsets[f * n1 + n1 - 1]; }
    sym_base[total] = acc;
    if (needed) *needed = acc;
    if (!symbols || sym_capacity < acc) { set_err(ctx, "symbol buffer too small"); return SO_E_NOMEM; }
    for (size_t f = 0; f < total; ++f) {
        const uint64_t n = sym_base[f + 1] - sym_base[f];
        if (n) CU(cudaMemcpyAsync(symbols + sym_base[f], ctx->sym_data + f * ctx->sym_frame_stride, n * sizeof(int16_t), cudaMemcpyDeviceToHost, st));
    }
    CU(cudaStreamSynchronize(st));
    return SO_OK;
}

// ---------------------------------------------------------------------------------------------------------
// decoder: restates decoder.decode (decoder.py:487-545) on packed arrays, host buffers in and out
// ---------------------------------------------------------------------------------------------------------
extern "C" int so_decode_sequence(so_ctx* ctx, const uint8_t* frame_types, const uint8_t* split, const int16_t* mv, const int16_t* levels,
                                  const int32_t* qp_rows_per_frame, int n_frames, int reset_at_intra, uint8_t* out_frames) {
    if (!ctx || !frame_types || !split || !mv || !levels || !out_frames || n_frames < 1) return SO_E_INVALID;
    // the arrays are indexed with the CONTEXT's geometry (n_frames x n_blocks of ctx): reject contents that cannot come from
    // a stream of this geometry before anything is enqueued (reference indices select ring slots on the device)
    for (int f = 0; f < n_frames; ++f) {
        if (frame_types[f] > 1) { set_err(ctx, "frame type must be 0 or 1"); return SO_E_INVALID; }
        const uint8_t* sp = split + (size_t)f * ctx->nblk;
        const int16_t* m = mv + (size_t)f * ctx->nblk * 12;
        for (int b = 0; b < ctx->nblk; ++b) {
            if (sp[b] > 1) { set_err(ctx, "split flag must be 0 or 1"); return SO_E_INVALID; }
            if (frame_types[f] == 1)
                for (int k = 0; k < 4; ++k)
                    if (m[(size_t)b * 12 + k * 3 + 2] < 0 || m[(size_t)b * 12 + k * 3 + 2] >= ctx->p.n_ref_frames) {
                        set_err(ctx, "reference index outside nRefFrames of the context"); return SO_E_INVALID;
                    }
        }
    }
    CU(cudaSetDevice(ctx->device));
    int rc = ensure_seq(ctx, (size_t)n_frames);
    if (rc) return rc;
    cudaStream_t st = ctx->stream;
    const FrameGeom& g = ctx->g;
    const size_t px = ctx->frame_px;
    CU(cudaMemcpyAsync(ctx->sq_split, split, (size_t)n_frames * ctx->nblk, cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(ctx->sq_mv, mv, (size_t)n_frames * ctx->nblk * 12 * sizeof(int16_t), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(ctx->sq_levels, levels, (size_t)n_frames * px * sizeof(int16_t), cudaMemcpyHostToDevice, st));
    rc = so_ref_reset(ctx, SO_ALL_UNITS, st);
    if (rc) return rc;
    const int nt = nthreads_px(g.bs);
    for (int f = 0; f < n_frames; ++f) {
        so_frame_out o;
        o.split = ctx->sq_split + (size_t)f * ctx->nblk;
        o.mv = ctx->sq_mv + (size_t)f * ctx->nblk * 12;
        o.levels = ctx->sq_levels + (size_t)f * px;
        o.recon = ctx->sq_recon + (size_t)f * px;
        o.row_sizes = ctx->sq_rows;
        o.stats = ctx->sq_stats;
        if (qp_rows_per_frame) CU(cudaMemcpyAsync(ctx->qp_rows_dev, qp_rows_per_frame + (size_t)f * g.nby, sizeof(int) * g.nby, cudaMemcpyHostToDevice, st));
        // decoder.py:504-509 decodes EVERY frame of a ParallelMode-1 stream as inter, so it cannot read the I frames a scene
        // cut (RCFlag 2, Encoder.py:1851-1856) puts into such a stream; with the encoder's list semantics
        // (reset_at_intra == 0, the round-trip mode) the frame type is honoured instead
        const bool intra = frame_types[f] == 0 && (ctx->p.parallel_mode != 1 || !reset_at_intra);
        if (!intra && ctx->p.parallel_mode == 1) { rc = ref_reset_impl(ctx, st, false); if (rc) return rc; }     // decoder.py:504-509
        if (!intra) {
            if (ctx->rs().list.empty()) { set_err(ctx, "inter frame with an empty reference list"); return SO_E_STATE; }
            rc = ensure_planes(ctx, 1, st);
            if (rc) return rc;
        }
        FlowArgs a = make_flow(ctx, o.recon, px, &o, n_frames, 0, ctx->p.qp);
        a.qp_rows = qp_rows_per_frame ? ctx->qp_rows_dev : nullptr;
        a.qp_blocks = (ctx->qp_blocks_dev && f < ctx->qp_blocks_frames) ? ctx->qp_blocks_dev + (size_t)f * ctx->nblk : nullptr;
        dim3 grid(ctx->nblk, 1);
        if (g.bs == 16) decode_block_kernel<16><<<grid, nt, 0, st>>>(a, intra ? 1 : 0);
        else if (g.bs == 8) decode_block_kernel<8><<<grid, nt, 0, st>>>(a, intra ? 1 : 0);
        else decode_block_kernel<4><<<grid, nt, 0, st>>>(a, intra ? 1 : 0);
        if (intra) {
            dim3 grid2(g.nby, 1);
            if (g.bs == 16 && g.W % 16 == 0 && std::getenv("SO_INTRA_GENERIC") == nullptr) intra_recon16_kernel<<<grid2, 256, (size_t)g.nbx * 9 + 16, st>>>(a);
            else if (g.bs == 16) intra_recon_kernel<16><<<grid2, nt, 0, st>>>(a);
            else if (g.bs == 8) intra_recon_kernel<8><<<grid2, nt, 0, st>>>(a);
            else intra_recon_kernel<4><<<grid2, nt, 0, st>>>(a);
            if (reset_at_intra) ctx->rs().list.clear();                                     // decoder.py:520 `ref_frames = []`
        }
        CU(cudaGetLastError());
        if (f < n_frames - 1 && ctx->p.parallel_mode != 1) {
            rc = ring_push(ctx, o.recon, px, 1, st);
            if (rc) return rc;
        }
    }
    CU(cudaMemcpyAsync(out_frames, ctx->sq_recon, (size_t)n_frames * px, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return SO_OK;
}

extern "C" int so_last_timing(so_ctx* ctx, double out[4]) {
    if (!ctx || !out) return SO_E_INVALID;
    if (ctx->timing_pending) {
        CU(cudaSetDevice(ctx->device));
        CU(cudaEventSynchronize(ctx->ev1));
        float ms = 0;
        cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
        ctx->timing[0] = ms;
        double me = 0, tq = 0;
        for (auto& p : ctx->ev_me) { float t = 0; cudaEventElapsedTime(&t, p.first, p.second); me += t; }
        for (auto& p : ctx->ev_tq) { float t = 0; cudaEventElapsedTime(&t, p.first, p.second); tq += t; }
        ctx->timing[1] = me; ctx->timing[2] = tq; ctx->timing[3] = (double)ctx->launches;
        ctx->timing[4] = (double)ctx->ev_me.size();
        double xs = 0;
        for (size_t i : ctx->ev_xs) { float t = 0; cudaEventElapsedTime(&t, ctx->ev_me[i].first, ctx->ev_me[i].second); xs += t; }
        ctx->timing[5] = xs; ctx->timing[6] = (double)ctx->ev_xs.size();
        double fi = 0;
        for (size_t i : ctx->ev_tq_inter) { float t = 0; cudaEventElapsedTime(&t, ctx->ev_tq[i].first, ctx->ev_tq[i].second); fi += t; }
        ctx->timing[7] = fi; ctx->timing[8] = (double)ctx->ev_tq_inter.size();
        ctx->timing_pending = false;
    }
    for (int i = 0; i < 4; ++i) out[i] = ctx->timing[i];
    return SO_OK;
}

// number of exhaustive-search kernel launches covered by timing [1] of the last run
extern "C" int so_last_me_launches(so_ctx* ctx) {
    double t[4];
    if (!ctx) return SO_E_INVALID;
    int rc = so_last_timing(ctx, t);
    return rc ? rc : (int)ctx->timing[4];
}

// the inter finish kernel (transform / quantisation / RLE size / reconstruction of P frames) alone in the last so_seq_run:
// out[0] = summed CUDA-event time (ms) of its timed launches, out[1] = timed launches (one per timed P frame, all units in
// one launch), out[2] = the same sum over all transform kernels incl. intra frames, out[3] = timed frames
extern "C" int so_last_finish_timing(so_ctx* ctx, double out[4]) {
    double t[4];
    if (!ctx || !out) return SO_E_INVALID;
    int rc = so_last_timing(ctx, t);
    if (rc) return rc;
    out[0] = ctx->timing[7]; out[1] = ctx->timing[8]; out[2] = ctx->timing[2]; out[3] = (double)ctx->timed_frames;
    return SO_OK;
}

// exhaustive-search kernels alone (me_ring_kernel / me_tma_kernel): out[0] = summed CUDA-event time (ms) of the timed
// launches, out[1] = timed launches, out[2] = frames with per-kernel events, out[3] = frames of the run
extern "C" int so_last_search_timing(so_ctx* ctx, double out[4]) {
    double t[4];
    if (!ctx || !out) return SO_E_INVALID;
    int rc = so_last_timing(ctx, t);
    if (rc) return rc;
    out[0] = ctx->timing[5]; out[1] = ctx->timing[6]; out[2] = (double)ctx->timed_frames; out[3] = (double)ctx->total_frames;
    return SO_OK;
}

// ---------------------------------------------------------------------------------------------------------
// host text formatters (Encoder.py:1419-1542, canonical integers)
// ---------------------------------------------------------------------------------------------------------
namespace {
struct Out {
    char* dst; int64_t cap; int64_t n = 0;
    void ch(char c) { if (n < cap) dst[n] = c; ++n; }
    void str(const char* s) { while (*s) ch(*s++); }
    void num(long v) {
        char buf[24]; int k = 0;
        unsigned long u = v < 0 ? (unsigned long)(-v) : (unsigned long)v;
        do { buf[k++] = (char)('0' + u % 10); u /= 10; } while (u);
        if (v < 0) ch('-');
        while (k) ch(buf[--k]);
    }
    void tuple3(long a, long b, long c) { ch('('); num(a); str(", "); num(b); str(", "); num(c); ch(')'); }
    int64_t finish() { if (n < cap) dst[n] = 0; return n < cap ? n : -(n + 1); }
};
}  // namespace

extern "C" int64_t so_format_mv_frame(int frame_type, const uint8_t* split, const int16_t* mv, int n_blocks, int blocks_per_row,
                                      const int32_t* qp_rows, char* dst, int64_t cap) {
    Out o{dst, cap};
    o.num(frame_type); o.ch('|');
    long ref[3] = {0, 0, 0};
    long ref_qp = 0;
    for (int b = 0; b < n_blocks; ++b) {
        if (b) o.ch(';');
        if (qp_rows && b % blocks_per_row == 0) {
            const long q = qp_rows[b / blocks_per_row];
            o.num(q - ref_qp); o.ch('@');
            ref_qp = q;
        }
        const int16_t* m = mv + (size_t)b * 12;
        if (frame_type == 0) {
            if (!split[b]) { o.str("0'("); o.num(m[0] - ref[0]); o.ch(')'); ref[0] = m[0]; }
            else {
                o.str("1'(");
                for (int k = 0; k < 4; ++k) { if (k) o.ch(','); o.num(m[k * 3] - ref[0]); ref[0] = m[k * 3]; }
                o.ch(')');
            }
        } else {
            if (!split[b]) {
                o.str("0'"); o.tuple3(m[0] - ref[0], m[1] - ref[1], m[2] - ref[2]);
                ref[0] = m[0]; ref[1] = m[1]; ref[2] = m[2];
            } else {
                o.str("1'(");
                for (int k = 0; k < 4; ++k) {
                    if (k) o.ch(',');
                    o.tuple3(m[k * 3] - ref[0], m[k * 3 + 1] - ref[1], m[k * 3 + 2] - ref[2]);
                    ref[0] = m[k * 3]; ref[1] = m[k * 3 + 1]; ref[2] = m[k * 3 + 2];
                }
                o.ch(')');
            }
        }
    }
    return o.finish();
}

static void rle_block_text(Out& o, const int16_t* lev, int pitch, int n) {
    // entropy_encoder_block (Encoder.py:1086-1131) printed as a Python list
    o.ch('[');
    bool first = true;
    auto emit = [&](long v) { if (!first) o.str(", "); first = false; o.num(v); };
    int16_t vals[256];
    int nz = 0; long zeros = 0; int flag = 1;
    for (int d = 0; d < 2 * n - 1; ++d) {
        int i = d < n ? 0 : d - n + 1, j = d < n ? d : n - 1;
        for (; i < n && j >= 0; ++i, --j) {
            const int16_t v = lev[(size_t)i * pitch + j];
            if (v != 0) {
                if (flag == 0) { if (zeros) { emit(zeros); zeros = 0; } nz = 0; flag = 1; }
                vals[nz++] = v;
            } else {
                if (flag == 1) { if (nz) { emit(-nz); for (int q = 0; q < nz; ++q) emit(vals[q]); nz = 0; } zeros = 0; flag = 0; }
                ++zeros;
            }
        }
    }
    if (nz) { emit(-nz); for (int q = 0; q < nz; ++q) emit(vals[q]); }
    if (zeros) emit(0);
    o.ch(']');
}

extern "C" int64_t so_format_residual_frame(const uint8_t* split, const int16_t* levels, int width, int height, int block_size,
                                            char* dst, int64_t cap) {
    Out o{dst, cap};
    const int nbx = width / block_size, nby = height / block_size, S = block_size / 2;
    for (int b = 0; b < nbx * nby; ++b) {
        if (b) o.ch(';');
        const int x = (b % nbx) * block_size, y = (b / nbx) * block_size;
        if (!split[b]) {
            o.str("0'(");
            rle_block_text(o, levels + (size_t)y * width + x, width, block_size);
            o.ch(')');
        } else {
            o.str("1'(");
            for (int k = 0; k < 4; ++k) {
                if (k) o.ch(',');
                rle_block_text(o, levels + (size_t)(y + (k >> 1) * S) * width + x + (k & 1) * S, width, S);
            }
            o.ch(')');
        }
    }
    return o.finish();
}

// Whole-sequence text bitstream (transmit_bitstream, Encoder.py:1544-1573 with the parseable residual format): every frame
// is formatted by the per-frame formatters above on a pool of host threads, the two files are written in frame order.
// qp_rows_per_frame i32 [n_frames][height / block_size] or NULL (RCFlag off).  n_threads <= 0: hardware concurrency.
static int write_bitstream_impl(const uint8_t* frame_types, const uint8_t* split, const int16_t* mv, const int16_t* levels,
                                const int16_t* symbols, const uint64_t* sym_pos, const uint32_t* sym_count,
                                const int32_t* qp_rows_per_frame, int n_frames, int width, int height, int block_size,
                                const char* mv_path, const char* residual_path, int n_threads) {
    if (!frame_types || !split || !mv || !mv_path || !residual_path || n_frames < 1 || block_size < 2 ||
        width % block_size || height % block_size) return SO_E_INVALID;
    const int nbx = width / block_size, nby = height / block_size, nblk = nbx * nby;
    const size_t px = (size_t)width * height;
    int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    nt = std::max(1, std::min(nt, std::min(n_frames, 64)));
    FILE* fm = fopen(mv_path, "wb");
    FILE* fr = fopen(residual_path, "wb");
    if (!fm || !fr) { if (fm) fclose(fm); if (fr) fclose(fr); return SO_E_INVALID; }
    auto fmt = [](auto&& call, size_t guess, std::string& out) {
        out.resize(guess);
        int64_t n = call(&out[0], (int64_t)out.size());
        if (n == INT64_MIN) return false;
        if (n < 0) { out.resize((size_t)(-n) + 16); n = call(&out[0], (int64_t)out.size()); }
        out.resize((size_t)std::max<int64_t>(n, 0));
        return true;
    };
    bool ok = true;
    // batches of nt frames: format in parallel, then write in order (bounds the memory held as text)
    std::vector<std::string> mvt(nt), rst(nt);
    std::vector<char> good(nt, 1);
    for (int f0 = 0; f0 < n_frames && ok; f0 += nt) {
        const int nb = std::min(nt, n_frames - f0);
        std::vector<std::thread> th;
        for (int i = 0; i < nb; ++i) {
            th.emplace_back([&, i]() {
                const int f = f0 + i;
                const uint8_t* sp = split + (size_t)f * nblk;
                const int16_t* m = mv + (size_t)f * nblk * 12;
                const int32_t* qp = qp_rows_per_frame ? qp_rows_per_frame + (size_t)f * nby : nullptr;
                bool g = fmt([&](char* d, int64_t c) { return so_format_mv_frame(frame_types[f], sp, m, nblk, nbx, qp, d, c); }, 64 + (size_t)nblk * 48, mvt[i]);
                if (levels) {
                    const int16_t* lv = levels + (size_t)f * px;
                    g = fmt([&](char* d, int64_t c) { return so_format_residual_frame(sp, lv, width, height, block_size, d, c); }, 1024 + px / 2, rst[i]) && g;
                } else {
                    const int16_t* sy = symbols + sym_pos[f];
                    const int64_t ns = sym_count[f];
                    g = fmt([&](char* d, int64_t c) { return so_format_residual_frame_packed(sp, sy, ns, nblk, block_size, d, c); },
                            1024 + (size_t)ns * 6 + (size_t)nblk * 16, rst[i]) && g;
                }
                good[i] = g ? 1 : 0;
            });
        }
        for (auto& t : th) t.join();
        for (int i = 0; i < nb; ++i) {
            ok = ok && good[i];
            mvt[i].push_back('\n'); rst[i].push_back('\n');
            ok = ok && fwrite(mvt[i].data(), 1, mvt[i].size(), fm) == mvt[i].size();
            ok = ok && fwrite(rst[i].data(), 1, rst[i].size(), fr) == rst[i].size();
        }
    }
    ok = (fclose(fm) == 0) && ok;
    ok = (fclose(fr) == 0) && ok;
    return ok ? SO_OK : SO_E_INVALID;
}

extern "C" int so_write_bitstream_files(const uint8_t* frame_types, const uint8_t* split, const int16_t* mv, const int16_t* levels,
                                        const int32_t* qp_rows_per_frame, int n_frames, int width, int height, int block_size,
                                        const char* mv_path, const char* residual_path, int n_threads) {
    if (!levels) return SO_E_INVALID;
    return write_bitstream_impl(frame_types, split, mv, levels, nullptr, nullptr, nullptr, qp_rows_per_frame, n_frames, width, height,
                                block_size, mv_path, residual_path, n_threads);
}

// the same two files from packed symbol streams (so_set_symbol_output): no levels needed on the host
extern "C" int so_write_bitstream_files_symbols(const uint8_t* frame_types, const uint8_t* split, const int16_t* mv, const int16_t* symbols,
                                                const uint64_t* sym_pos, const uint32_t* sym_count, const int32_t* qp_rows_per_frame,
                                                int n_frames, int width, int height, int block_size, const char* mv_path,
                                                const char* residual_path, int n_threads) {
    if (!sym_pos || !sym_count) return SO_E_INVALID;
    return write_bitstream_impl(frame_types, split, mv, nullptr, symbols, sym_pos, sym_count, qp_rows_per_frame, n_frames, width, height,
                                block_size, mv_path, residual_path, n_threads);
}

// Parser of the two text streams (decode_differential_entropy, decoder.py:590-690): the inverse of
// so_write_bitstream_files.  Frames are independent (the differential vectors restart at every frame), so the lines are
// parsed in parallel on host threads.  Outputs as so_encode_sequence packs them; qp_rows i32 [n_frames][height /
// block_size] is filled when rc_on != 0 (may be NULL otherwise).  Returns SO_E_INVALID on malformed or short input.
namespace {
struct Cursor {
    const char* p; const char* e;
    bool bad = false;                            // a number with too many digits was met: the line is malformed
    bool number(long& v) {                       // next integer at or after p (skips anything that is not a digit or '-')
        while (p < e && !((*p >= '0' && *p <= '9') || (*p == '-' && p + 1 < e && p[1] >= '0' && p[1] <= '9'))) ++p;
        if (p >= e) return false;
        bool neg = false;
        if (*p == '-') { neg = true; ++p; }
        long x = 0;
        int digits = 0;
        while (p < e && *p >= '0' && *p <= '9') {
            if (++digits > 9) { bad = true; return false; }      // nothing in the format exceeds 9 digits: no overflow
            x = x * 10 + (*p - '0'); ++p;
        }
        v = neg ? -x : x;
        return true;
    }
};

static void scan_order_tab(int n, std::vector<int>& idx) {      // Encoder.py:1095-1123 / decoder.py:573-586
    idx.clear();
    for (int k = 0; k < 2 * n - 1; ++k) {
        int i = k < n ? 0 : k - n + 1, j = k < n ? k : n - 1;
        while (i < n && j >= 0) { idx.push_back(i * n + j); ++i; --j; }
    }
}

static bool parse_mv_line_c(const char* b, const char* e, int nblk, int bpr, bool rc_on, uint8_t* ftype, uint8_t* split, int16_t* mv,
                            int32_t* qp_rows) {
    Cursor c{b, e};
    long t;
    if (!c.number(t) || c.p >= e || *c.p != '|' || (t != 0 && t != 1)) return false;      // frame type: 0 intra, 1 inter
    ++c.p;
    *ftype = (uint8_t)t;
    long ref[3] = {0, 0, 0}, ref_qp = 0;
    memset(split, 0, (size_t)nblk);
    memset(mv, 0, (size_t)nblk * 12 * sizeof(int16_t));
    for (int j = 0; j < nblk; ++j) {
        const char* ie = (const char*)memchr(c.p, ';', (size_t)(e - c.p));
        if (!ie) ie = e;
        Cursor it{c.p, ie};
        if (rc_on && j % bpr == 0) {
            const char* at = (const char*)memchr(it.p, '@', (size_t)(ie - it.p));
            long q;
            Cursor qc{it.p, at ? at : ie};
            if (!at || !qc.number(q)) return false;
            ref_qp += q;
            if (ref_qp < 0 || ref_qp > 15) return false;
            if (qp_rows) qp_rows[j / bpr] = (int32_t)ref_qp;
            it.p = at + 1;
        }
        if (it.p >= ie || (*it.p != '0' && *it.p != '1') || it.p + 1 >= ie || it.p[1] != '\'') return false;
        const bool sp = *it.p == '1';
        it.p += 2;
        split[j] = sp ? 1 : 0;
        const int nvec = sp ? 4 : 1;
        for (int k = 0; k < nvec; ++k) {
            long v[3] = {0, 0, 0};
            const int nc = t == 0 ? 1 : 3;
            for (int q = 0; q < nc; ++q) if (!it.number(v[q])) return false;
            // accumulated vectors must fit the packed int16 fields (and a reference index its range)
            if (t == 0) {
                ref[0] += v[0];
                if (ref[0] < -32768 || ref[0] > 32767) return false;
                mv[(size_t)j * 12 + k * 3] = (int16_t)ref[0];
            } else {
                for (int q = 0; q < 3; ++q) {
                    ref[q] += v[q];
                    if (ref[q] < -32768 || ref[q] > 32767) return false;
                    mv[(size_t)j * 12 + k * 3 + q] = (int16_t)ref[q];
                }
                if (ref[2] < 0 || ref[2] >= SO_MAX_REF) return false;
            }
        }
        c.p = ie < e ? ie + 1 : e;
        if (j < nblk - 1 && ie >= e) return false;
    }
    return true;
}

static bool parse_res_line_c(const char* b, const char* e, int W, int H, int bs, const uint8_t* split, int16_t* lev,
                             const std::vector<int>& o_full, const std::vector<int>& o_sub) {
    const int nbx = W / bs, nblk = nbx * (H / bs), sub = bs / 2;
    memset(lev, 0, (size_t)W * H * sizeof(int16_t));
    const char* p = b;
    for (int blk = 0; blk < nblk; ++blk) {
        const char* ie = (const char*)memchr(p, ';', (size_t)(e - p));
        if (!ie) ie = e;
        if (p + 1 >= ie || p[1] != '\'' || (p[0] - '0') != split[blk]) return false;
        const int y = (blk / nbx) * bs, x = (blk % nbx) * bs;
        const int nl = split[blk] ? 4 : 1, n = split[blk] ? sub : bs;
        const std::vector<int>& ord = split[blk] ? o_sub : o_full;
        const char* q = p + 2;
        for (int k = 0; k < nl; ++k) {
            const char* lb = (const char*)memchr(q, '[', (size_t)(ie - q));
            if (!lb) return false;
            const char* le = (const char*)memchr(lb, ']', (size_t)(ie - lb));
            if (!le) return false;
            int16_t* dst = lev + (size_t)(y + (split[blk] ? (k >> 1) * sub : 0)) * W + x + (split[blk] ? (k & 1) * sub : 0);
            Cursor c{lb + 1, le};
            int pos = 0;
            long s;
            while (c.number(s)) {                  // entropy_decoder_block, decoder.py:548-586
                if (s < 0) {
                    if (-s > (long)(n * n - pos)) return false;         // a run cannot be longer than what is left of the block
                    for (long i = 0; i < -s; ++i) {
                        long v;
                        if (!c.number(v) || v < -32768 || v > 32767) return false;
                        const int o = ord[pos++];
                        dst[(size_t)(o / n) * W + o % n] = (int16_t)v;
                    }
                } else if (s == 0) break;
                else {
                    if (s > (long)(n * n - pos)) return false;
                    pos += (int)s;
                }
            }
            if (c.bad) return false;
            q = le + 1;
        }
        p = ie < e ? ie + 1 : e;
        if (blk < nblk - 1 && ie >= e) return false;
    }
    return true;
}

static bool read_file(const char* path, std::string& out) {
    FILE* f = fopen(path, "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    const long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    out.resize((size_t)std::max(0L, n));
    const bool ok = n <= 0 || fread(&out[0], 1, (size_t)n, f) == (size_t)n;
    fclose(f);
    return ok;
}

static void line_spans(const std::string& s, std::vector<std::pair<const char*, const char*>>& out) {
    const char* p = s.data();
    const char* e = p + s.size();
    while (p < e) {
        const char* nl = (const char*)memchr(p, '\n', (size_t)(e - p));
        const char* le = nl ? nl : e;
        bool blank = true;
        for (const char* q = p; q < le; ++q) if (*q != ' ' && *q != '\r' && *q != '\t') { blank = false; break; }
        if (!blank) out.emplace_back(p, le);
        p = nl ? nl + 1 : e;
    }
}
}  // namespace

extern "C" int so_parse_bitstream_files(const char* mv_path, const char* residual_path, int n_frames, int width, int height, int block_size,
                                        int rc_on, uint8_t* frame_types, uint8_t* split, int16_t* mv, int16_t* levels, int32_t* qp_rows,
                                        int n_threads) {
    if (!mv_path || !residual_path || !frame_types || !split || !mv || !levels || n_frames < 1 || block_size < 2 ||
        width % block_size || height % block_size || (rc_on && !qp_rows)) return SO_E_INVALID;
    std::string mvs, rss;
    if (!read_file(mv_path, mvs) || !read_file(residual_path, rss)) return SO_E_INVALID;
    std::vector<std::pair<const char*, const char*>> ml, rl;
    line_spans(mvs, ml); line_spans(rss, rl);
    if ((int)ml.size() < n_frames || (int)rl.size() < n_frames) return SO_E_INVALID;
    const int nbx = width / block_size, nby = height / block_size, nblk = nbx * nby;
    const size_t px = (size_t)width * height;
    std::vector<int> o_full, o_sub;
    scan_order_tab(block_size, o_full); scan_order_tab(block_size / 2, o_sub);
    int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    nt = std::max(1, std::min(nt, std::min(n_frames, 64)));
    std::vector<int> okv(nt, 1);
    std::vector<std::thread> th;
    for (int w = 0; w < nt; ++w) {
        th.emplace_back([&, w]() {
            for (int f = w; f < n_frames; f += nt) {
                uint8_t* sp = split + (size_t)f * nblk;
                bool ok = parse_mv_line_c(ml[f].first, ml[f].second, nblk, nbx, rc_on != 0, frame_types + f, sp, mv + (size_t)f * nblk * 12,
                                          rc_on ? qp_rows + (size_t)f * nby : nullptr);
                ok = ok && parse_res_line_c(rl[f].first, rl[f].second, width, height, block_size, sp, levels + (size_t)f * px, o_full, o_sub);
                if (!ok) okv[w] = 0;
            }
        });
    }
    for (auto& t : th) t.join();
    for (int v : okv) if (!v) return SO_E_INVALID;
    return SO_OK;
}

extern "C" int64_t so_format_residual_frame_symbols(const uint8_t* split, const uint32_t* offsets, const int16_t* symbols, int n_blocks,
                                                    char* dst, int64_t cap) {
    Out o{dst, cap};
    for (int b = 0; b < n_blocks; ++b) {
        if (b) o.ch(';');
        const int nseg = split[b] ? 4 : 1;
        o.str(split[b] ? "1'(" : "0'(");
        for (int k = 0; k < nseg; ++k) {
            if (k) o.ch(',');
            o.ch('[');
            const uint32_t s0 = offsets[(size_t)b * 4 + k], s1 = offsets[(size_t)b * 4 + k + 1];
            for (uint32_t i = s0; i < s1; ++i) { if (i > s0) o.str(", "); o.num(symbols[i]); }
            o.ch(']');
        }
        o.ch(')');
    }
    return o.finish();
}

// ---------------------------------------------------------------------------------------------------------
// packed symbol streams on the host (what sequence encodes deliver with so_set_symbol_output): a frame is the
// concatenation of its blocks' run-level lists in raster order, the four lists of a split block in Z order.  The lists
// are self-delimiting (entropy_decoder_block, decoder.py:548-586): a list ends with the symbol 0 (trailing zeros) or
// when its runs have covered all n*n coefficients.
// ---------------------------------------------------------------------------------------------------------
namespace {
// walk one list: calls val(position in scan order, value) for every non-zero coefficient; returns the number of symbols
// consumed, or -1 when the stream is inconsistent with an n x n block
template <typename V>
static int64_t walk_list(const int16_t* sy, int64_t avail, int n, V&& val) {
    const int nn = n * n;
    int pos = 0;
    int64_t i = 0;
    while (pos < nn) {
        if (i >= avail) return -1;
        const int s = sy[i++];
        if (s == 0) return i;                           // trailing zeros
        if (s > 0) { if (s >= nn - pos) return -1; pos += s; continue; }      // a zero run never reaches the end (that is the 0 symbol)
        const int cnt = -s;
        if (cnt > nn - pos || i + cnt > avail) return -1;
        for (int k = 0; k < cnt; ++k) { if (sy[i + k] == 0) return -1; val(pos + k, sy[i + k]); }
        i += cnt; pos += cnt;
    }
    return i;
}
}  // namespace

// residual text of one frame (entropy_encoder_frame, Encoder.py:1522-1542) from its packed symbols; same bytes as
// so_format_residual_frame.  Returns bytes written, -(needed + 1) when cap is too small, or INT64_MIN on a corrupt stream.
extern "C" int64_t so_format_residual_frame_packed(const uint8_t* split, const int16_t* symbols, int64_t n_symbols, int n_blocks,
                                                   int block_size, char* dst, int64_t cap) {
    Out o{dst, cap};
    int64_t at = 0;
    for (int b = 0; b < n_blocks; ++b) {
        if (b) o.ch(';');
        const int nseg = split[b] ? 4 : 1, n = split[b] ? block_size / 2 : block_size;
        o.str(split[b] ? "1'(" : "0'(");
        for (int k = 0; k < nseg; ++k) {
            if (k) o.ch(',');
            const int64_t used = walk_list(symbols + at, n_symbols - at, n, [](int, int) {});
            if (used < 0) return INT64_MIN;
            o.ch('[');
            for (int64_t i = 0; i < used; ++i) { if (i) o.str(", "); o.num(symbols[at + i]); }
            o.ch(']');
            at += used;
        }
        o.ch(')');
    }
    if (at != n_symbols) return INT64_MIN;
    return o.finish();
}

// packed symbols -> quantised levels (the inverse of entropy_encoder_block over whole frames), frames in parallel on host
// threads.  levels i16 [n_frames][height][width] is fully overwritten.
extern "C" int so_symbols_to_levels(const uint8_t* split, const int16_t* symbols, const uint64_t* sym_pos, const uint32_t* sym_count,
                                    int n_frames, int width, int height, int block_size, int16_t* levels, int n_threads) {
    if (!split || !sym_pos || !sym_count || !levels || n_frames < 1 || block_size < 2 || width % block_size || height % block_size)
        return SO_E_INVALID;
    const int nbx = width / block_size, nblk = nbx * (height / block_size), sub = block_size / 2;
    const size_t px = (size_t)width * height;
    std::vector<int> o_full, o_sub;
    scan_order_tab(block_size, o_full); scan_order_tab(sub, o_sub);
    int nt = n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency();
    nt = std::max(1, std::min(nt, std::min(n_frames, 64)));
    std::vector<int> okv(nt, 1);
    auto work = [&](int w) {
        for (int f = w; f < n_frames; f += nt) {
            int16_t* lev = levels + (size_t)f * px;
            memset(lev, 0, px * sizeof(int16_t));
            const uint8_t* sp = split + (size_t)f * nblk;
            const int16_t* sy = symbols + sym_pos[f];
            const int64_t ns = sym_count[f];
            int64_t at = 0;
            bool ok = symbols != nullptr || ns == 0;
            for (int b = 0; b < nblk && ok; ++b) {
                const int y = (b / nbx) * block_size, x = (b % nbx) * block_size;
                const int nseg = sp[b] ? 4 : 1, n = sp[b] ? sub : block_size;
                const std::vector<int>& ord = sp[b] ? o_sub : o_full;
                for (int k = 0; k < nseg && ok; ++k) {
                    int16_t* dst = lev + (size_t)(y + (sp[b] ? (k >> 1) * sub : 0)) * width + x + (sp[b] ? (k & 1) * sub : 0);
                    const int64_t used = walk_list(sy + at, ns - at, n, [&](int p, int v) { const int o = ord[p]; dst[(size_t)(o / n) * width + o % n] = (int16_t)v; });
                    if (used < 0) ok = false; else at += used;
                }
            }
            if (!ok || at != ns) okv[w] = 0;
        }
    };
    if (nt == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int w = 0; w < nt; ++w) th.emplace_back(work, w);
        for (auto& t : th) t.join();
    }
    for (int v : okv) if (!v) return SO_E_INVALID;
    return SO_OK;
}
