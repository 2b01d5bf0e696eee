// Shared device-side definitions for the StreamOptima B200 kernels (sm_100a only).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define SO_MAX_REF 8

// Geometry + per-frame launch parameters, passed by value to every kernel.
struct FrameGeom {
    int W, H;          // coded luma size (multiples of bs)
    int pitch;         // byte pitch of the internal u8 planes (multiple of 16, >= W)
    int bs;            // block size
    int nbx, nby;      // blocks per row / column
    int r;             // integer search range
    int R;             // candidate range in search units: 2r when fme else r
    int fme;           // half-pel enabled
    int nref;          // number of frames currently in the reference list (1..SO_MAX_REF)
};

// Reference list of one unit as the kernels see it: nref entries, 4 phase planes each
// (0: integer, 1: horizontal half, 2: vertical half, 3: diagonal); non-FME uses plane 0 only.
struct RefList {
    const uint8_t* plane[SO_MAX_REF][4];
};

// Exact int <-> double conversions on the FP64 add pipe.  I2F.F64 / F2I.F64 run on the XU pipe at a small fraction of a
// lane per clock; with 64 of them per 16x16 block they -- not the transform -- bounded the finish kernels (ncu:
// sm__inst_executed_pipe_xu 89 %).  so_i2d: bits (0x43300000, v ^ 0x80000000) are the double 2^52 + 2^31 + v; so_d2i_rint:
// d + 1.5 * 2^52 rounds to nearest-even like rint() and leaves the integer in the low word (|d| < 2^31).
__device__ __forceinline__ double so_i2d(int v) {
    return __dadd_rn(__hiloint2double(0x43300000, v ^ (int)0x80000000), -4503601774854144.0);
}
__device__ __forceinline__ int so_d2i_rint(double d) {
    return __double2loint(__dadd_rn(d, 6755399441055744.0));
}

// The reference ring as the kernels address it: [unit][slot][phase 0..3][byte shift 0..3][H][pitch].  Shift plane c of
// a phase holds the phase plane moved left by c bytes (plane_c[x] = plane[x + c], zero past the frame edge): a TMA box
// load at a 16-byte aligned x of plane c delivers a search window whose candidates at x = c (mod 4) are word aligned.
struct RefRing {
    const uint8_t* base;
    size_t unit_stride, slot_stride, plane_stride;     // plane_stride: between phases (shift 0 of each)
    int slot[SO_MAX_REF];                              // list index -> ring slot
    __device__ __host__ const uint8_t* plane(int unit, int idx, int ph) const {
        return base + unit * unit_stride + slot[idx] * slot_stride + ph * plane_stride;
    }
};

// Result of a motion search for one (sub-)block.
struct __align__(8) MeResult {
    int16_t dx, dy, ref;
    int16_t none;      // 1: no valid candidate (MAE = inf, mv = fallback)
    uint32_t sad;      // SAD of the winner; for fast ME the reference returns the ref index as "MAE" (quirk Q4)
};

// Programmatic dependent launch: pdl_trigger() lets the next kernel of the stream be scheduled once every CTA of this grid
// has called it (or exited); pdl_wait() blocks until the previous kernel of the stream has completed and its writes are
// visible.  Both are no-ops for a grid launched without the attribute.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ uint32_t sad4_acc(uint32_t a, uint32_t b, uint32_t acc) {
    uint32_t d;
    asm("vabsdiff4.u32.u32.u32.add %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(acc));
    return d;   // SASS: VABSDIFF4.U8.ACC
}

// Sample of the (2H-1)x(2W-1) half-pel frame of Encoder.py:388-406 held as four phase planes (appendix A1).
__device__ __forceinline__ int up_at(const RefList& rl, int ref, int pitch, int X, int Y) {
    return rl.plane[ref][((Y & 1) << 1) | (X & 1)][(size_t)(Y >> 1) * pitch + (X >> 1)];
}

// Predictor sample (i,j) of a bs x bs block whose top-left is (px0,py0) in search units (half-pel when fme) displaced
// by mv.  Restates calculate_inter_frame_residual (Encoder.py:444-456) and reconstruct_frame (Encoder.py:862-873,
// 907-919).  second_bs < 0: residual-path test `p + 2*bs < size - bs`; otherwise the split-reconstruction test
// `p + second_bs < size - second_bs` with the parent size (quirk Q5, Encoder.py:910).
struct PredSel {
    int mode;      // 0: in-bounds block, 1: constant 128, 2: zero-padded contiguous crop
    int PX, PY;
};

__device__ __forceinline__ PredSel pred_select(const FrameGeom& g, int x0, int y0, int dx, int dy, int bs, int second_bs) {
    PredSel s;
    const int Wr = g.fme ? 2 * g.W - 1 : g.W;
    const int Hr = g.fme ? 2 * g.H - 1 : g.H;
    s.PX = x0 + dx;
    s.PY = y0 + dy;
    if (s.PX >= 0 && s.PX < Wr - bs && s.PY >= 0 && s.PY < Hr - bs) {
        s.mode = 0;
        if (g.fme) {
            bool ok;
            if (second_bs < 0) ok = (s.PX + 2 * bs < Wr - bs) && (s.PY + 2 * bs < Hr - bs);
            else ok = (s.PX + second_bs < Wr - second_bs) && (s.PY + second_bs < Hr - second_bs);
            if (!ok) s.mode = 1;
        }
    } else {
        s.mode = 2;
    }
    return s;
}

__device__ __forceinline__ int pred_sample(const FrameGeom& g, const RefList& rl, const PredSel& s, int ref, int i, int j) {
    if (s.mode == 1) return 128;
    if (g.fme) {
        if (s.mode == 0) return up_at(rl, ref, g.pitch, s.PX + 2 * i, s.PY + 2 * j);
        const int X = s.PX + i, Y = s.PY + j;       // handle_boundary_conditions: contiguous crop of the half-pel frame
        if (X < 0 || Y < 0 || X >= 2 * g.W - 1 || Y >= 2 * g.H - 1) return 0;
        return up_at(rl, ref, g.pitch, X, Y);
    }
    const int X = s.PX + i, Y = s.PY + j;
    if (s.mode == 2 && (X < 0 || Y < 0 || X >= g.W || Y >= g.H)) return 0;
    return rl.plane[ref][0][(size_t)Y * g.pitch + X];
}

// Candidate validity rectangle of find_best_match / fast_motion_estimation (Encoder.py:695-698, 728-730) in search
// units for a block at pixel position p (one axis): lo <= d <= hi.  `second` applies the `+2*bs` test.
__device__ __forceinline__ void valid_range(int p, int size_px, int bs, int fme, int second, int& lo, int& hi) {
    const int size = fme ? 2 * size_px - 1 : size_px;
    const int p0 = fme ? 2 * p : p;
    lo = -p0;
    hi = size - bs - 1 - p0;
    if (second) hi = min(hi, size - 3 * bs - 1 - p0);
}

// Quantiser (appendix A2): level = round_half_even(tc / 2^s) in integers; s = qp + (0|1|2) by anti-diagonal.
__device__ __forceinline__ int q_shift(int u, int v, int n, int qp) {
    const int d = u + v;
    return qp + (d < n - 1 ? 0 : (d == n - 1 ? 1 : 2));
}
__device__ __forceinline__ int quant_rhe(int tc, int s) {
    if (s == 0) return tc;
    const int q = tc >> s;                       // floor
    const int rem = tc - (q << s);
    const int half = 1 << (s - 1);
    return q + (rem > half ? 1 : (rem == half ? (q & 1) : 0));
}
