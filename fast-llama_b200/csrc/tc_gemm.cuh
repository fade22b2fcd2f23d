// tc_gemm.cuh — group-scaled INT8 GEMM on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// quant::matmul<int8_t> for SEVERAL activation rows (quant_operators.cpp:252-284, driven by forward() with bs > 1,
// transformer.cpp:105-151):    out[i][j] = sum over groups g, ascending:  fma(ws[j][g] * xs[i][g], float(dot_g(W[j], X[i])), acc)
// The integer dot of one group is exact whatever computes it; the FP32 chain over groups is the part whose order fixes
// the bits.  So: one `tcgen05.mma.kind::i8` pair per 64-wide group (one MMA for the 32-wide groups of Q8_0) produces the
// group's INT32 dots for 128 weight rows x N activation rows into a TMEM accumulator, and the epilogue warps walk the
// chain in registers, group after group, exactly as the reference does.  Bit-identical to the GEMV path by construction.
//
// Shape: weights are the M = 128 operand (A), activation rows the N operand (B), both K-major, no swizzle: the canonical
// layout of 8-row x 16-byte core matrices,  byte(row r, k) = (k / 16) * LBO + (r / 8) * 128 + (r % 8) * 16 + k % 16  with
// LBO = 16 * rows.  Weights are packed ONCE at upload into exactly that image, stage after stage in the order a CTA
// consumes them, so a stage is one `cp.async.bulk` (no tensor maps), HBM bytes == algorithmic bytes, and a CTA may own any
// number of rows (<= 128 per tile): all 148 SMs stream every matrix, as in the persistent decode kernel.
//
// Roles (320 threads): warp 0 = bulk-copy producer (weights never depend on activations: under programmatic dependent
// launch it fills the ring while the previous kernel is still running, and only then waits for the activations);
// warp 1 = MMA issuer (one thread) + TMEM owner; warps 2-9 = two epilogue warpgroups.  TMEM holds up to 8 accumulator
// "units" of a few groups each (one full/empty barrier pair per unit, not per group: the issuer thread and the epilogue
// warps pay their synchronisation latency once per unit), so the MMAs of the next units run while a unit's chains are walked.
//   W1/W3 (DUAL): the two matrices are separate MMAs of one stage into adjacent TMEM columns of the same lanes; warpgroup 0
//   owns the W1 chains, warpgroup 1 the W3 chains, SwiGLU joins them through shared memory.
//   otherwise the two warpgroups split the activation rows (columns) of one accumulator.
#pragma once
#include "megakernel.cuh"

namespace fl {

constexpr int kTcKC = 256;               // bytes (= int8 elements) of K per stage
constexpr int kTcThreads = 320;
constexpr int kTcMaxRows = 128;          // rows per tile = MMA M
constexpr int kTcMaxSlots = 12;
constexpr int kTcXchgCols = 16;          // SwiGLU exchange: columns per piece ([16][128] floats = 8 KB)

struct TcPart { int rb, nr, nt; };
__host__ __device__ inline TcPart tc_part(int M, int c, int G) {
    TcPart r;
    r.rb = (int)((long long)M * c / G);
    r.nr = (int)((long long)M * (c + 1) / G) - r.rb;
    r.nt = (r.nr + kTcMaxRows - 1) / kTcMaxRows;
    return r;
}
__host__ __device__ inline void tc_tile(const TcPart& pt, int t, int& lr0, int& R) {
    lr0 = pt.nr * t / pt.nt;
    R = pt.nr * (t + 1) / pt.nt - lr0;
}
// bytes of one weight stage: tt images of kTcKC/16 x R pieces of 16 bytes, then tt x gps x R group scales
__host__ __device__ inline int tc_stage_a_bytes(int R, int tt, int gs) {
    return tt * kTcKC * R + ((tt * (kTcKC / gs) * R * 4 + 15) & ~15);
}
// activation image: per K chunk  [kTcKC/16][N] pieces, then [gps][N] scales
__host__ __device__ inline int tc_chunk_b_bytes(int N, int gs) { return kTcKC * N + (kTcKC / gs) * N * 4; }
__host__ __device__ inline int tc_pad_n(int T) { return T <= 8 ? 8 : T <= 16 ? 16 : T <= 32 ? 32 : 64; }
__host__ __device__ inline size_t tc_image_bytes(int K, int N, int gs) { return (size_t)((K + kTcKC - 1) / kTcKC) * tc_chunk_b_bytes(N, gs); }
// byte offset of element k of activation row i inside the image, and of the scale of its group g
__host__ __device__ inline size_t tc_xq_offset(int i, int k, int N, int gs) {
    return (size_t)(k / kTcKC) * tc_chunk_b_bytes(N, gs) + (size_t)((k % kTcKC) / 16) * N * 16 + (size_t)i * 16 + k % 16;
}
__host__ __device__ inline size_t tc_xs_offset(int i, int g, int N, int gs) {
    const int gps = kTcKC / gs;
    return (size_t)(g / gps) * tc_chunk_b_bytes(N, gs) + (size_t)kTcKC * N + (size_t)((g % gps) * N + i) * 4;
}

// ---------------------------------------------------------------------------------------------- packing
// rows [row_base, row_base + rows_src) of the logical matrix (M_total rows) from the reference's row-major int8 payload +
// scale table into the stage stream; sub-stream m of tt (W1 = 0 / W3 = 1 of the fused matrix).  grid (n_ctas, y).
template <int GS>
__global__ void pack_tc_kernel(const uint8_t* __restrict__ raw, const float* __restrict__ scales, uint8_t* __restrict__ packed,
                               const unsigned long long* __restrict__ cta_off, int M_total, int K, int row_base, int rows_src, int tt, int m) {
    constexpr int GPS = kTcKC / GS;
    const int c = blockIdx.x, G = gridDim.x;
    const TcPart pt = tc_part(M_total, c, G);
    const int nkc = (K + kTcKC - 1) / kTcKC, Gtot = K / GS;
    uint8_t* tile_base = packed + cta_off[c];
    for (int t = 0; t < pt.nt; ++t) {
        int lr0, R;
        tc_tile(pt, t, lr0, R);
        const int sb = tc_stage_a_bytes(R, tt, GS);
        for (int kc = blockIdx.y; kc < nkc; kc += gridDim.y) {
            uint8_t* st = tile_base + (size_t)kc * sb;
            const int n_pieces = (kTcKC / 16) * R;
            for (int idx = threadIdx.x; idx < n_pieces + GPS * R; idx += blockDim.x) {
                if (idx < n_pieces) {
                    const int k16 = idx / R, r = idx - k16 * R;
                    const int srow = pt.rb + lr0 + r - row_base;
                    if (srow < 0 || srow >= rows_src) continue;
                    const int k = kc * kTcKC + k16 * 16;
                    uint4 v = make_uint4(0, 0, 0, 0);
                    if (k < K) v = *reinterpret_cast<const uint4*>(raw + (size_t)srow * K + k);
                    *reinterpret_cast<uint4*>(st + (size_t)m * kTcKC * R + (size_t)idx * 16) = v;
                } else {
                    const int si = idx - n_pieces, g = si / R, r = si - g * R;
                    const int srow = pt.rb + lr0 + r - row_base;
                    if (srow < 0 || srow >= rows_src) continue;
                    const int gi = kc * GPS + g;
                    *reinterpret_cast<float*>(st + (size_t)tt * kTcKC * R + ((size_t)(m * GPS + g) * R + r) * 4) = gi < Gtot ? scales[(size_t)srow * Gtot + gi] : 0.0f;
                }
            }
        }
        tile_base += (size_t)nkc * sb;
    }
}

// natural-order quantised rows [T][K] + scales [T][K/GS] -> activation image (tests / per-op entry point)
template <int GS>
__global__ void tc_pack_x_kernel(const int8_t* __restrict__ xq, const float* __restrict__ xs, uint8_t* __restrict__ img, int T, int K, int N) {
    const int G = K / GS;
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < (size_t)T * K; idx += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(idx / K), k = (int)(idx % K);
        img[tc_xq_offset(i, k, N, GS)] = (uint8_t)xq[idx];
    }
    for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < (size_t)T * G; idx += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(idx / G), g = (int)(idx % G);
        *reinterpret_cast<float*>(img + tc_xs_offset(i, g, N, GS)) = xs[idx];
    }
}

// ---------------------------------------------------------------------------------------------- tcgen05 PTX
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_alloc(uint32_t* smem_slot, uint32_t ncols) {      // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(smem_slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_dealloc(uint32_t taddr, uint32_t ncols) {         // the same warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(ncols) : "memory");
}
// all tcgen05 operations this thread issued so far -> one arrival on `bar` when they have completed
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
// shared-memory matrix descriptor, K-major, no swizzle: start address, leading (K) and stride (M/N) byte offsets in 16-byte
// units, descriptor version 1 (Blackwell), base offset 0, layout type 0 = SWIZZLE_NONE
__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3ffffu) >> 4) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) | ((uint64_t)((sbo >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
// instruction descriptor: D = S32 (bits 4-5 = 2), A = B = signed int8 (bits 7-9, 10-12 = 1), both K-major (bits 15, 16 = 0),
// N >> 3 in bits 17-22, M >> 4 in bits 24-28; dense, no saturation
__host__ __device__ constexpr uint32_t tc_idesc_i8(int M, int N) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// descriptors travel as (lo, hi) halves: hi (stride offset, version) is constant, lo = start address | leading offset << 16
// Called by ALL lanes of the issuer warp with warp-uniform operands; one elected lane issues.  The election lives inside
// the asm so that the surrounding control flow stays warp-uniform: with the whole loop under `if (lane == 0)` (round 2, first
// version) every operand was a per-lane value and the compiler wrapped each UTCIMMA in R2UR moves and an ELECT / BRA.U.ANY
// waterfall loop - 116 cycles per MMA, 128-344 MMAs per CTA per matrix, i.e. most of a GEMM's run time at N = 8
// (profiles/r02/ncu_full_r02_qgemm_decode8.csv: t = 12 us + 0.059 us x MMAs for all four matrices, tensor pipe 0.15 % active).
template <bool ACC>
__device__ __forceinline__ void tc_mma_i8(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc) {
    asm volatile(
        "{\n\t.reg .pred p, e;\n\t.reg .b64 da, db;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
        "@e tcgen05.mma.cta_group::1.kind::i8 [%0], da, db, %4, {%6, %6, %6, %6}, p;\n\t}"
        :: "r"(tmem_d), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "n"(ACC ? 1 : 0), "r"(0u) : "memory");
}
// all tcgen05 operations the elected lane issued so far -> one arrival on `bar` (all lanes call, one commits)
__device__ __forceinline__ void tc_commit_elect(uint64_t* bar) {
    asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
                 "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, int (&d)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7]),
                   "=r"(d[8]), "=r"(d[9]), "=r"(d[10]), "=r"(d[11]), "=r"(d[12]), "=r"(d[13]), "=r"(d[14]), "=r"(d[15])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_ld8(uint32_t taddr, int (&d)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_ld4(uint32_t taddr, int (&d)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// programmatic dependent launch
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------- the kernel
enum { TC_EPI_STORE = 0, TC_EPI_RESADD = 1, TC_EPI_SWIGLU = 2 };

struct TcGemmArgs {
    const uint8_t* w;                      // packed stage stream of the matrix (W1/W3: the fused stream)
    const unsigned long long* cta_off;     // [gridDim + 1] byte offsets of the CTAs' streams
    const uint8_t* xq;                     // activation image, N rows wide
    float* out;                            // [T][ldo] fp32 (STORE: written; RESADD: out += ; SWIGLU: hd)
    int M;                                 // output rows (SWIGLU: hidden)
    int K;
    int T;                                 // live activation rows (<= N)
    int ldo;
    int smem_bytes;                        // dynamic shared memory handed to the launch
    int variant;                           // reserved (0)
};

template <int GS, int N, bool DUAL>
struct TcShape {
    static constexpr int TT = DUAL ? 2 : 1;
    static constexpr int GPS = kTcKC / GS;                 // groups per stage
    static constexpr int MPG = GS / 32;                    // K = 32 MMAs per group
    static constexpr int NB = TT * N;                      // TMEM columns per group
    static constexpr int GU = (128 / NB) >= GPS ? GPS : ((128 / NB) >= 1 ? (128 / NB) : 1);     // groups per TMEM unit (<= 128 columns)
    static constexpr int UPS = GPS / GU;                   // units per stage
    static constexpr int NUNITS = (512 / (GU * NB)) > 8 ? 8 : (512 / (GU * NB));
    // the two epilogue warpgroups share the columns of one accumulator.  Also at N = 8 (4 columns each): the epilogue is one
    // latency-bound instruction stream per live TMEM quadrant (28 rows = one warp), and it - not the MMA issue (halving the
    // MMAs: 3.05 -> 2.89 ms per step) nor the weight stream - paces the thin GEMMs at ~0.65 us per stage
    static constexpr bool SPLIT = !DUAL;
    static constexpr int NC = (DUAL || !SPLIT) ? N : N / 2;   // columns (activation rows) per epilogue thread
    static constexpr int EPW = (DUAL || SPLIT) ? 8 : 4;    // epilogue warps that take part
    static_assert(GPS % GU == 0 && NUNITS >= 2, "unit geometry");
};

template <int GS, int N, bool DUAL, int EPI>
__global__ void __launch_bounds__(kTcThreads, 1) qgemm_kernel(const __grid_constant__ TcGemmArgs a) {
    using S = TcShape<GS, N, DUAL>;
    constexpr int TT = S::TT, GPS = S::GPS, MPG = S::MPG, NB = S::NB, GU = S::GU, UPS = S::UPS, NUNITS = S::NUNITS, NC = S::NC, EPW = S::EPW;
    constexpr bool SPLIT = S::SPLIT;
    static_assert(!DUAL || EPI == TC_EPI_SWIGLU, "the fused W1/W3 stream ends in SwiGLU");
    static_assert(N == 8 || N == 16 || N == 32 || N == 64, "N");

    extern __shared__ __align__(128) uint8_t tc_smem[];
    uint8_t* smem = tc_smem;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);          // [kTcMaxSlots]
    uint64_t* empty = full + kTcMaxSlots;                         // [kTcMaxSlots]
    uint64_t* tfull = empty + kTcMaxSlots;                        // [8]
    uint64_t* tempty = tfull + 8;                                 // [8]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 8);
    float* xchg = reinterpret_cast<float*>(smem + 512);           // [kTcXchgCols][128] (DUAL only)
    uint8_t* ring = smem + 512 + (DUAL ? kTcXchgCols * 128 * 4 : 0);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const TcPart pt = tc_part(a.M, blockIdx.x, gridDim.x);
    const int nkc = (a.K + kTcKC - 1) / kTcKC, Gtot = a.K / GS;
    constexpr int b_bytes = kTcKC * N + GPS * N * 4;
    int r_max = 0;
    for (int t = 0; t < pt.nt; ++t) { int lr0, R; tc_tile(pt, t, lr0, R); r_max = R > r_max ? R : r_max; }
    const int slot_bytes = (b_bytes + tc_stage_a_bytes(r_max, TT, GS) + 127) & ~127;
    int n_slots = (a.smem_bytes - (int)(ring - smem) - 2048) / slot_bytes;        // 2 KB tail: an M = 128 MMA reads past a short tile's rows
    if (n_slots > kTcMaxSlots) n_slots = kTcMaxSlots;
    const int total_stages = pt.nt * nkc;

    if (tid == 0) {
        for (int i = 0; i < n_slots; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1 + EPW); }
        for (int i = 0; i < NUNITS; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], EPW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tc_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
    pdl_launch_dependents();

    if (warp == 0) {
        // ================= producer =================
        if (lane == 0 && total_stages > 0) {
            const uint8_t* wsrc = a.w + a.cta_off[blockIdx.x];
            // phase 1: weights of the first ring-full of stages (they do not depend on the previous kernel)
            const int pre = total_stages < n_slots ? total_stages : n_slots;
            {
                const uint8_t* src = wsrc;
                int t = 0, kc = 0, lr0, R;
                tc_tile(pt, 0, lr0, R);
                for (int i = 0; i < pre; ++i) {
                    const uint32_t ab = (uint32_t)tc_stage_a_bytes(R, TT, GS);
                    mbar_arrive_expect_tx(&full[i], ab + (uint32_t)b_bytes);
                    bulk_g2s(ring + (size_t)i * slot_bytes + b_bytes, src, ab, &full[i]);
                    src += ab;
                    if (++kc == nkc) { kc = 0; if (++t < pt.nt) tc_tile(pt, t, lr0, R); }
                }
            }
            pdl_wait();                       // the activation image is complete and visible
            {
                int kc = 0;
                for (int i = 0; i < pre; ++i) {
                    bulk_g2s(ring + (size_t)i * slot_bytes, a.xq + (size_t)kc * b_bytes, (uint32_t)b_bytes, &full[i]);
                    if (++kc == nkc) kc = 0;
                }
            }
            // phase 2: steady state
            uint32_t slot = 0, par = 0;       // stage `pre` reuses slot 0 after its first release
            const uint8_t* src = wsrc;
            int i = 0;
            for (int t = 0; t < pt.nt; ++t) {
                int lr0, R;
                tc_tile(pt, t, lr0, R);
                const uint32_t ab = (uint32_t)tc_stage_a_bytes(R, TT, GS);
                for (int kc = 0; kc < nkc; ++kc, ++i, src += ab) {
                    if (i < pre) continue;
                    mbar_wait_sleep(&empty[slot], par);
                    mbar_arrive_expect_tx(&full[slot], ab + (uint32_t)b_bytes);
                    uint8_t* dst = ring + (size_t)slot * slot_bytes;
                    bulk_g2s(dst + b_bytes, src, ab, &full[slot]);
                    bulk_g2s(dst, a.xq + (size_t)kc * b_bytes, (uint32_t)b_bytes, &full[slot]);
                    if (++slot == (uint32_t)n_slots) { slot = 0; par ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        {
            // M = 64 where the CTA's tiles have at most 64 rows (Wo, W2: 28-35 rows per CTA): the MMA reads half the A rows from
            // shared memory.  Measured: no change (8 sequences 3.01 vs 2.96 ms per step, profiles/r02/rows_bench_v3.log) - the
            // ~95 cycles per K = 32 MMA are not the A read
            const uint32_t idesc = tc_idesc_i8(r_max <= 64 ? 64 : 128, N);
            const uint32_t desc_hi = (128u >> 4) | (1u << 14);       // stride (M/N) byte offset 128, descriptor version 1
            uint32_t slot = 0, par = 0, un = 0, upar = 1;            // tempty parity 1 passes on a fresh barrier
            for (int t = 0; t < pt.nt; ++t) {
                int lr0, R;
                tc_tile(pt, t, lr0, R);
                const uint32_t lbo_a = (uint32_t)R, lbo_b = (uint32_t)N;     // leading (K) byte offsets / 16
                for (int kc = 0; kc < nkc; ++kc) {
                    mbar_wait(&full[slot], par);
                    tc_fence_after();
                    const uint32_t b_img = (smem_u32(ring + (size_t)slot * slot_bytes) & 0x3ffffu) >> 4;     // in 16-byte units
                    const uint32_t a_img = b_img + (uint32_t)(b_bytes >> 4);
                    const int ng = (Gtot - kc * GPS) < GPS ? (Gtot - kc * GPS) : GPS;
#pragma unroll
                    for (int u = 0; u < UPS; ++u) {
                        if (u * GU < ng) {
                            mbar_wait(&tempty[un], upar);
                            tc_fence_after();
#pragma unroll
                            for (int gg = 0; gg < GU; ++gg) {
                                const int g = u * GU + gg;
                                if (g < ng) {
#pragma unroll
                                    for (int m = 0; m < TT; ++m) {
                                        const uint32_t dcol = tmem_base + un * (GU * NB) + (uint32_t)(gg * NB + m * N);
#pragma unroll
                                        for (int k32 = 0; k32 < MPG; ++k32) {
                                            const uint32_t k16 = (uint32_t)(g * (GS / 16) + k32 * 2);
                                            const uint32_t a_lo = (a_img + (uint32_t)m * (kTcKC / 16) * lbo_a + k16 * lbo_a) | (lbo_a << 16);
                                            const uint32_t b_lo = (b_img + k16 * lbo_b) | (lbo_b << 16);
                                            if (k32 == 0) tc_mma_i8<false>(dcol, a_lo, b_lo, desc_hi, idesc);
#ifndef FL_TC_SKIP_HALF        // timing experiment (wrong results): is the GEMM bound by the MMA issue?
                                            else tc_mma_i8<true>(dcol, a_lo, b_lo, desc_hi, idesc);
#endif
                                        }
                                    }
                                }
                            }
                            tc_commit_elect(&tfull[un]);
                            if (++un == (uint32_t)NUNITS) { un = 0; upar ^= 1u; }
                        }
                    }
                    tc_commit_elect(&empty[slot]);      // the stage's operands have been read once every MMA above has completed
                    if (++slot == (uint32_t)n_slots) { slot = 0; par ^= 1u; }
                }
            }
        }
    } else if (warp - 2 < EPW) {
        // ================= epilogue: the FP32 chain over groups, one weight row per thread =================
        const int ew = warp - 2, wg = ew >> 2;
        const int q = warp & 3;                                  // the TMEM lane quadrant this warp may read
        // accumulator row -> TMEM lane: M = 128: row i in lane i; M = 64: row i in lane (i % 16) + 32 * (i / 16), i.e. the first
        // 16 lanes of every quadrant (cute/atom/mma_traits_sm100.hpp, half-subpartition atom)
        const int rpq = r_max <= 64 ? 16 : 32;                   // rows per quadrant
        const int r = q * rpq + lane;                            // local row inside the tile
        const int c0 = DUAL ? 0 : (SPLIT ? wg * NC : 0);         // first activation row of this thread
        const int msel = DUAL ? wg : 0;                          // which matrix of the fused stream
        const uint32_t tcol = (uint32_t)(DUAL ? wg * N : c0);
        pdl_wait();                                              // RESADD reads `out`; every kernel of the chain waits
        uint32_t slot = 0, par = 0, un = 0, fpar = 0;
        for (int t = 0; t < pt.nt; ++t) {
            int lr0, R;
            tc_tile(pt, t, lr0, R);
            const bool quad_live = q * rpq < R;
            const bool live = lane < rpq && r < R;
            const int row = pt.rb + lr0 + r;
            float acc[NC];
#pragma unroll
            for (int i = 0; i < NC; ++i) acc[i] = 0.0f;
            for (int kc = 0; kc < nkc; ++kc) {
                mbar_wait(&full[slot], par);
                const uint8_t* st = ring + (size_t)slot * slot_bytes;
                const float* xs_st = reinterpret_cast<const float*>(st + kTcKC * N);                              // [GPS][N]
                const float* ws_st = reinterpret_cast<const float*>(st + b_bytes + TT * kTcKC * R) + msel * GPS * R; // [GPS][R]
                const int ng = (Gtot - kc * GPS) < GPS ? (Gtot - kc * GPS) : GPS;
#pragma unroll
                for (int u = 0; u < UPS; ++u) {
                    if (u * GU < ng) {
                        mbar_wait(&tfull[un], fpar);
                        tc_fence_after();
                        if (quad_live) {
                            const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + un * (GU * NB) + tcol;
                            if constexpr (NC == 4) {
                                int d[GU][4];
#pragma unroll
                                for (int gg = 0; gg < GU; ++gg) tc_ld4(tbase + (uint32_t)(gg * NB), d[gg]);
                                tc_ld_wait();
#pragma unroll
                                for (int gg = 0; gg < GU; ++gg) {
                                    const int g = u * GU + gg;
                                    if (g < ng) {
                                        const float ws = live ? ws_st[g * R + r] : 0.0f;
                                        const float4 x0 = *reinterpret_cast<const float4*>(xs_st + g * N + c0);
                                        const float xs[4] = {x0.x, x0.y, x0.z, x0.w};
#pragma unroll
                                        for (int i = 0; i < 4; ++i) acc[i] = __fmaf_rn(__fmul_rn(ws, xs[i]), __int2float_rn(d[gg][i]), acc[i]);
                                    }
                                }
                            } else if constexpr (NC == 8) {
                                // all groups of the unit at once: GU x 8 columns, NB columns apart
                                int d[GU][8];
#pragma unroll
                                for (int gg = 0; gg < GU; ++gg) tc_ld8(tbase + (uint32_t)(gg * NB), d[gg]);
                                tc_ld_wait();
#pragma unroll
                                for (int gg = 0; gg < GU; ++gg) {
                                    const int g = u * GU + gg;
                                    if (g < ng) {
                                        const float ws = live ? ws_st[g * R + r] : 0.0f;
                                        const float4 x0 = *reinterpret_cast<const float4*>(xs_st + g * N + c0), x1 = *reinterpret_cast<const float4*>(xs_st + g * N + c0 + 4);
                                        const float xs[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
                                        for (int i = 0; i < 8; ++i) acc[i] = __fmaf_rn(__fmul_rn(ws, xs[i]), __int2float_rn(d[gg][i]), acc[i]);
                                    }
                                }
                            } else {
#pragma unroll
                                for (int gg = 0; gg < GU; ++gg) {
                                    const int g = u * GU + gg;
                                    if (g < ng) {
                                        const float ws = live ? ws_st[g * R + r] : 0.0f;
#pragma unroll
                                        for (int cb = 0; cb < NC; cb += 16) {
                                            int d[16];
                                            tc_ld16(tbase + (uint32_t)(gg * NB + cb), d);
                                            tc_ld_wait();
#pragma unroll
                                            for (int i4 = 0; i4 < 16; i4 += 4) {
                                                const float4 x = *reinterpret_cast<const float4*>(xs_st + g * N + c0 + cb + i4);
                                                acc[cb + i4] = __fmaf_rn(__fmul_rn(ws, x.x), __int2float_rn(d[i4]), acc[cb + i4]);
                                                acc[cb + i4 + 1] = __fmaf_rn(__fmul_rn(ws, x.y), __int2float_rn(d[i4 + 1]), acc[cb + i4 + 1]);
                                                acc[cb + i4 + 2] = __fmaf_rn(__fmul_rn(ws, x.z), __int2float_rn(d[i4 + 2]), acc[cb + i4 + 2]);
                                                acc[cb + i4 + 3] = __fmaf_rn(__fmul_rn(ws, x.w), __int2float_rn(d[i4 + 3]), acc[cb + i4 + 3]);
                                            }
                                        }
                                    }
                                }
                            }
                        }
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tempty[un]);          // the unit's values are in registers / consumed
                        if (++un == (uint32_t)NUNITS) { un = 0; fpar ^= 1u; }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[slot]);              // the stage's scales have been read
                if (++slot == (uint32_t)n_slots) { slot = 0; par ^= 1u; }
            }
            // ---- the tile's rows are complete
            if (EPI == TC_EPI_SWIGLU) {
                // simd::swiglu (x86_simd.cpp:1766-1770): warpgroup 1 hands its W3 rows over, kTcXchgCols activation rows at a time
#pragma unroll
                for (int cb = 0; cb < NC; cb += kTcXchgCols) {
                    constexpr int PC = NC < kTcXchgCols ? NC : kTcXchgCols;
                    if (wg == 1) {
#pragma unroll
                        for (int i = 0; i < PC; ++i) xchg[i * 128 + r] = acc[cb + i];
                    }
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                    if (wg == 0 && live) {
#pragma unroll
                        for (int i = 0; i < PC; ++i)
                            if (cb + i < a.T) a.out[(size_t)(cb + i) * a.ldo + row] = swiglu_exact(acc[cb + i], xchg[i * 128 + r]);
                    }
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                }
            } else if (live) {
#pragma unroll
                for (int i = 0; i < NC; ++i) {
                    if (c0 + i < a.T) {
                        float* o = a.out + (size_t)(c0 + i) * a.ldo + row;
                        if (EPI == TC_EPI_STORE) *o = acc[i];
                        else *o = __fadd_rn(*o, acc[i]);                // x1 += tmp (tensor.cpp:723)
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tc_dealloc(tmem_base, 512);
}

}  // namespace fl
