// megakernel.cuh — the decode step as ONE persistent sm_100a kernel (one CTA per SM, no grid barriers).
//
// Why: a decode token is 161 dependent matrix-vector phases of 18-140 MB each.  Launched as separate kernels every
// phase pays launch latency, a cold pipeline and its activation prologue with HBM idle.  Here the weight stream never
// stops:
//
//   * warp 8 (one elected lane) is a TMA producer: it walks the token's static weight schedule and issues
//     cp.async.bulk global->shared copies into a ring of stages guarded by full/empty mbarriers.  Weights do not
//     depend on activations, so the producer runs ahead across phase, layer and token boundaries; only ring capacity
//     (~150 KB per SM = ~3.4 us of this SM's HBM share) limits it.  A serial section shorter than that costs nothing;
//     a longer one cannot be caught up on (one SM ingests at most ~54 GB/s, and L2 bandwidth is no higher than HBM's).
//   * warps 0-7 are consumers: per phase they (1) fetch the phase's input vector, (2) rebuild the quantised activation
//     image in shared memory (rmsnorm chain, quantise), (3) drain their stages into (scale product, dot) pairs;
//     warp 9, the chain warp, turns the pairs into rows and publishes them (see the layout paragraph).
//   * CTAs exchange activations as TAGGED WORDS (value, tag) written and read with single 8-byte accesses, tag = a
//     number unique to (launch, token, layer, producer phase).  A reader issues all its loads at once and re-issues only
//     those whose tag is stale, so "wait for every producer" and "fetch the vector" are ONE L2 round trip instead of
//     barrier (two dependent round trips) + load (a third); profiles/r01/sync_bench_ABD.log has the measurements.
//     Write-after-read safety needs no extra synchronisation: a buffer is only overwritten by a phase that (transitively)
//     consumed data from every CTA produced AFTER that CTA's last read of the buffer (DESIGN.md "Tagged exchange").
//   * attention runs between the QKV and Wo phases on n_heads * CPH CTAs (CPH CTAs share one head: keys are split for
//     QK^T, head dims are split for the PV chains), exchanging the score vector through L2 the same way.
//
// Weight layout ("row per lane", packed once at upload): CTA c owns rows [M*c/G, M*(c+1)/G) of every matrix, cut into
// tiles of <= 32 rows, ONE LANE PER ROW.  A stage is 256 bytes of every row of a tile plus the rows' group scales, stored
// piece-major so lane i's 16-byte piece sits next to lane i+1's (conflict-free LDS.128), and a CTA's stages lie in HBM in
// exactly the order its producer issues them - its share of a phase is one contiguous byte range.
// The 8 consumer warps share the K chunks of a tile round-robin: for its chunk a warp computes, for every row (lane), the
// integer group dots (exact in any order) and the scale products, and drops the (product, float(dot)) pairs into a
// shared-memory pair buffer.  The reference's FP32 chain over groups (quant_operators.cpp:274) is strictly sequential, so
// a tenth warp - the CHAIN WARP - owns it: per superblock of 64 groups it waits on a named barrier for the 8 consumers,
// walks  acc = fma(product, dot, acc)  for its 32 rows (4.5 cycles per group), and after the last superblock applies the
// epilogue (store / residual add / SwiGLU / argmax) and publishes the rows.  Producer -> consumers -> chain warp is a
// three-stage pipeline: the expensive part (loads + dp4a) never waits for the serial part.
// (History, profiles/r01: a 4-rows-x-8-lanes layout that transposed pairs per stage ran at 1500 cycles per 8.7 KB stage per
// warp - slower than HBM; one warp per tile over the whole K starved on ILP; passing the chain as a token from warp to warp
// cost 8 hand-offs of ~0.2 us per tile on the critical path.)
//
// Hardware facts that shape the code (all measured, see DESIGN.md "What the profiler taught us"):
//   1. With ~227 KB of shared memory carved out there is practically no L1 left: every local-memory (stack) access and
//      every re-read of a global word is an L2 round trip.  Nothing here may spill or take the address of a local.
//   2. The kernel body must stay small (instruction cache: 130 -> 96 KB of code was worth 4 %): ONE instance of each phase
//      routine inside a flat, rolled loop; profiling counters and the event log are compiled in only with -DFL_PROFILE / -DFL_EVLOG.
//   3. An L2 round trip costs 0.3 us on an idle chip and 1-1.5 us while the weight stream saturates HBM; dependent round
//      trips are what a serial section is made of, so every one of them is counted.
//   4. A consumer may only wait on a ring slot whose stage has already been issued: mbarrier waits are by phase PARITY, and
//      a wait for revolution k+1 on a slot still filling for revolution k succeeds spuriously (found with
//      profiles/micro/sync_bench.cu).  The producer publishes its issue count; consumers check it first.
#pragma once
#include "kernels.cuh"

namespace fl {

constexpr int kConsumerWarps = 8;
constexpr int kConsumerThreads = kConsumerWarps * 32;
#ifdef FL_NO_SETMAXNREG
constexpr int kMegaThreads = kConsumerThreads + 64;     // A/B build: no register re-division (168 per thread everywhere)
#else
constexpr int kMegaThreads = kConsumerThreads + 128;    // + a third warpgroup: TMA producer warp, chain warp, two warps that exit at once
#endif
// Registers are re-divided between the warpgroups at kernel start (setmaxnreg): with three warps per scheduler the launch
// allocation is 168 per thread, which the consumers' phase loop does not fit without spilling (and with ~227 KB of shared
// memory carved out a spill is an L2 round trip); the producer and chain warps need far less.  3 x 168 = 88 + 2 x 208 per scheduler (with fewer than 88 ptxas serialises the chain warp's load batches).
constexpr int kRegsService = 88, kRegsConsumer = 208;
static_assert(kRegsService + 2 * kRegsConsumer <= 3 * 168, "the CTA's register pool is what the launch allocated: 3 warps x 168 per scheduler (a larger sum deadlocks in setmaxnreg.inc)");
constexpr int kPairGroups = 64;                          // groups (all sub-streams together) per superblock = per pair buffer (16 KB)
constexpr int kTagsPerLayer = 8;
constexpr int kVChunkRows = 32;                         // cached V rows per chunk of the attention part's ring (8 quads of 4 positions)
constexpr int kSerialWarp = kConsumerWarps - 1;   // runs the single-warp serial sections: the scheduler favours the highest warp id of a
                                                  // sub-partition, and warp 7 shares its sub-partition only with warp 3 (not with the producer / chain warps)
constexpr int kProfThread = kConsumerThreads - 1;  // keeps the profiling clock: last lane of the serial warp (in the PV group, not a chain lane)

struct MegaLayer {
    const uint8_t* qkv;
    const uint8_t* wo;
    const uint8_t* w13;
    const uint8_t* w2;
    const float* att_norm;
    const float* ffn_norm;
};

struct MegaParams {
    const MegaLayer* layers;
    const uint8_t* cls;
    const float* out_norm;
    const float* att_norm; const float* ffn_norm;     // [n_layers][dim] each
    const float* emb;
    const unsigned long long *off_qkv, *off_wo, *off_w13, *off_w2, *off_cls;   // per-CTA stream offsets inside the packed matrices
    uint2* x1t; uint2* qkvt; uint2* attnt; uint2* hdt;   // tagged activation vectors: word i = (float bits, tag)
    uint2* hdqt;                      // the QUANTISED hd vector, tagged: per group GS * ES / 4 payload words, then the scale (build_hd)
    uint2* score_t;                   // [n_heads][score_stride] tagged raw scores exchanged between the CTAs of a head
    uint4* am;                        // [gridDim] argmax partials (value bits, tag, index, tag)
    float* logits;
    float* k_cache; float* v_cache;   // this sequence: [n_layers][n_kv_heads][max_seq][HS]
    const float* rope;
    SeqState* st;
    int* out_tokens; int out_cap; int* argmax_out;
    float* tap_norm;
    unsigned long long* prof;         // optional [gridDim][32] ns per category, see fl_profile_read
    unsigned long long* evlog;        // optional event log of the traced layer on CTAs 7 and gridDim-3: [2][4096] {count | (ns << 24 | warp << 20 | type << 12 | arg)}
    int dim, hidden, n_layers, n_heads, n_kv_heads, vocab, max_seq;
    int qkv_rows;
    int score_stride;
    float attn_scale;
    int n_steps;
    int cph;                          // CTAs per head (1, 2 or 4)
    int n_slots;                      // ring stages
    int tile_cap, slot_bytes;         // rows of the tallest tile; bytes of a ring slot = one stage of such a tile
    int debug_skip;                   // profiling experiments only (FL_DEBUG_SKIP): 1 = consumers release stages without computing
    int window;                       // max stages in flight (issued, not yet landed); >= n_slots: no limit
    uint32_t epoch;                   // tags of this launch are epoch + 1 ... epoch + n_steps * (n_layers + 1) * 8
    // dynamic shared memory carve-up (byte offsets)
    int off_ring, off_xq, off_xs, off_xt, off_chain, off_att, off_misc, off_bars, off_vstage, off_vbars, off_pairs, off_psrc, off_geom;
    int v_chunk_rows, n_vchunks;     // V ring of the attention part: n_vchunks chunks of v_chunk_rows rows x HS/cph floats
    // several sequences in one launch (fl_forward_batch): sequence s uses the exchange buffers at + s * xchg_stride bytes, the
    // caches at + s * cache_stride floats, st[s], out_tokens + s * out_cap, argmax_out[s].  Every phase is walked once per
    // sequence back to back, so one sequence's exchanges and serial sections hide behind the others' weight streaming, and
    // the second pass over a phase's weights is served from L2.
    int n_seqs;
    unsigned long long xchg_stride, cache_stride;
};

constexpr int kMaxSeqsPerLaunch = 16;      // per-sequence state lives in 64 spare words of the misc block

// the per-sequence buffers of a launch (MS = false: the single sequence the pointers in MegaParams already describe)
struct SeqView {
    uint2 *x1t, *qkvt, *attnt, *hdt, *hdqt, *score_t;
    uint4* am;
    float *kc, *vc;
};
template <bool MS>
__device__ __forceinline__ SeqView seq_view(const MegaParams& p, int s) {
    SeqView v{p.x1t, p.qkvt, p.attnt, p.hdt, p.hdqt, p.score_t, p.am, p.k_cache, p.v_cache};
    if (MS) {
        const unsigned long long o = p.xchg_stride * (unsigned long long)s;
        v.x1t = reinterpret_cast<uint2*>(reinterpret_cast<char*>(p.x1t) + o);
        v.qkvt = reinterpret_cast<uint2*>(reinterpret_cast<char*>(p.qkvt) + o);
        v.attnt = reinterpret_cast<uint2*>(reinterpret_cast<char*>(p.attnt) + o);
        v.hdt = reinterpret_cast<uint2*>(reinterpret_cast<char*>(p.hdt) + o);
        v.hdqt = reinterpret_cast<uint2*>(reinterpret_cast<char*>(p.hdqt) + o);
        v.score_t = reinterpret_cast<uint2*>(reinterpret_cast<char*>(p.score_t) + o);
        v.am = reinterpret_cast<uint4*>(reinterpret_cast<char*>(p.am) + o);
        v.kc = p.k_cache + p.cache_stride * (unsigned long long)s;
        v.vc = p.v_cache + p.cache_stride * (unsigned long long)s;
    }
    return v;
}

// ---------------------------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return done != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try(bar, parity)) { }
}
// Producer-side wait: back off instead of spinning, the ring holds microseconds of data.
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity) {
    while (!mbar_try(bar, parity)) __nanosleep(32);
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// One ring stage issued by the producer WARP: all lanes call with warp-uniform operands, one elected lane arms the barrier, starts
// the bulk copy and publishes the new issue count with a release store (ordered after the arming).  With the loop under
// `if (lane == 0)` the compiler wrapped every UBLKCP in five R2UR.BROADCAST moves and an ELECT / BRA.U.ANY waterfall loop and the
// count needed a MEMBAR.SC.CTA: ~180 cycles per stage, i.e. 2 us to request a ring refill after the gate opens.
__device__ __forceinline__ void issue_stage_elect(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint32_t* issued, uint32_t count) {
#ifndef FL_OLD_PRODUCER
    asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
                 "@e mbarrier.arrive.expect_tx.shared::cta.b64 _, [%3], %2;\n\t"
                 "@e cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n\t"
                 "@e st.release.cta.shared.u32 [%4], %5;\n\t}"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "r"(smem_u32(issued)), "r"(count) : "memory");
#endif
}
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" :: "n"(kConsumerThreads) : "memory"); }
__device__ __forceinline__ unsigned long long gtimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ uint32_t ld_shared_volatile_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ void st_shared_volatile_u32(uint32_t* p, uint32_t v) {
    asm volatile("st.volatile.shared.u32 [%0], %1;" :: "r"(smem_u32(p)), "r"(v) : "memory");
}

// ---- tagged words: one 8-byte (value, tag) pair per element; single-copy atomic, so no fences are involved.
__device__ __forceinline__ uint4 ld_tag2(const uint2* p) {        // two consecutive tagged words (16-byte aligned)
    uint4 v;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ uint2 ld_tag1(const uint2* p) {
    uint2 v;
    asm volatile("ld.relaxed.gpu.global.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_tag(uint2* p, float v, uint32_t tag) {
    asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1,%2};" :: "l"(p), "r"(__float_as_uint(v)), "r"(tag) : "memory");
}
__device__ __forceinline__ uint4 ld_relaxed_v4(const uint4* p) {
    uint4 v;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_relaxed_v4(uint4* p, uint4 v) {
    asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// where the rmsnorm rebuild's sum-of-squares chain runs (A/B builds): 0 = on the chain warp, x * w under it; 1 = on the chain warp,
// x * w in front of it; 2 = on consumer warp 7, x * w of the other warps under it
#ifndef FL_RB_MODE
#define FL_RB_MODE 0
#endif
constexpr int kRbMode = FL_RB_MODE;

// sensitivity experiments (profiles/r02/delay_sensitivity.log): -DFL_DELAY_AT=n spins ~1 us at point n, else nothing
template <int POINT>
__device__ __forceinline__ void delay_point() {
#ifdef FL_DELAY_AT
    if (POINT == FL_DELAY_AT) { const long long t = clock64(); while (clock64() - t < 2000) { } }
#endif
}

// per-CTA phase timing, only when MegaParams::prof is set; lives in registers.  Thread kProfThread keeps the clock.
struct Prof {
    unsigned long long* p; unsigned long long t0; int trace_slot;   // trace_slot >= 0: the current build records when its input was complete
    __device__ __forceinline__ void stop(int tid, int cat) {
#ifdef FL_PROFILE        // libfastllama_b200_prof.so only: ~30 call sites are 6 KB of code the instruction cache needs
        if (p && tid == kProfThread) { const unsigned long long t = (unsigned long long)clock64(); atomicAdd(p + cat, t - t0); t0 = t; }      // SM cycles: %globaltimer costs ~0.2 us a read
#endif
    }
    unsigned long long* ev;           // event log of this CTA while the traced layer runs, else NULL
    static __device__ __forceinline__ unsigned long long evtime() {
#ifdef FL_EVLOG_CLOCK      // SM cycles instead of %globaltimer (256 ns steps): events inside one CTA only
        return (unsigned long long)clock64() & 0xffffffffffull;
#else
        return gtimer();
#endif
    }
    unsigned int* evn;                // its event counter (shared memory: a global atomic with a return value costs a round trip)
    __device__ __forceinline__ void log(int lane, int warp, int type, int arg) {
#ifdef FL_EVLOG          // the per-warp event log is a debugging build: it costs ~10 KB of code the instruction cache needs
        if (ev && lane == 0) { const unsigned int i = atomicAdd(evn, 1u); if (i < 4095u) ev[1 + i] = (evtime() << 24) | ((unsigned long long)warp << 20) | ((unsigned long long)type << 12) | (unsigned long long)(arg & 0xfff); ev[0] = i + 1; }
#endif
    }
    // absolute timestamp of one event of the traced layer (slots 22..31): skew and latency of one exchange, see profiles/trace_layer.py
    __device__ __forceinline__ void mark(int tid, int slot, bool on) {
#ifdef FL_PROFILE
        if (p && on && tid == kProfThread) p[slot] = gtimer();
#endif
    }
};

// ---------------------------------------------------------------------------------------------- schedule
constexpr int kStageRowBytes = 256;        // bytes of one row in one stage
constexpr int kTileRows = 32;              // one lane per row: the most a tile can hold (an engine's tiles are capped at MegaParams::tile_cap <= 32)

template <int QT, int GS>
struct Rk {
    static constexpr int ES = (QT == Q_INT8) ? 1 : 2;            // element bytes
    static constexpr int EPS = kStageRowBytes / ES;               // elements of a row per stage
    static constexpr int GPS = EPS / GS;                          // quantisation groups of a row per stage
    static constexpr int PPG = GS * ES / 16;                      // 16-byte pieces per group
    static constexpr int PIECES = kStageRowBytes / 16;
    __host__ __device__ static constexpr int stage_bytes(int R) { return R * kStageRowBytes + ((R * GPS * 4 + 15) & ~15); }
    static_assert(GS == 64 || (GS == 32 && QT == Q_INT8), "group size");
};

// rows of CTA c and their cut into tiles (host and device agree on this arithmetic; the packer bakes it into the layout)
struct RkPart { int rb, nr, nt; };
// cap: the engine's tile height.  The ring's slots are as large as the largest stage, so the cap is the largest tile the four
// matrices of a layer need under the 32-lane limit (28 rows at the 7B shape: 23 slots instead of 20 in the same shared memory),
// and the classifier, whose rows per CTA would cut into taller tiles (31), is cut with the same cap.
__host__ __device__ inline RkPart rk_part(int M, int c, int G, int cap) {
    RkPart r;
    r.rb = (int)((long long)M * c / G);
    r.nr = (int)((long long)M * (c + 1) / G) - r.rb;
    r.nt = (r.nr + cap - 1) / cap;
    return r;
}
__host__ __device__ inline void rk_tile(const RkPart& pt, int t, int& lr0, int& R) {
    lr0 = pt.nr * t / pt.nt;
    R = pt.nr * (t + 1) / pt.nt - lr0;
}

// A tile's K chunks are cut into nsb superblocks of at most kPairGroups / tt groups per sub-stream (one pair buffer).
__host__ __device__ inline void rk_superblock(int nkc, int sk, int j, int& k0, int& S) { k0 = j * sk; S = nkc - k0 < sk ? nkc - k0 : sk; }   // sk = kPairGroups / tt / gps chunks, the last one shorter

// Pack rows [row_base, row_base + rows_src) of the logical matrix (M_total rows; `tt` sub-streams, this source is
// sub-stream `m`: W1 = 0 / W3 = 1 of the fused W13 matrix) from the reference's row-major payload + scale table.
// CTA stream order = issue order: [tile][K chunk][m]; stage = pieces p = 0..15 x R lanes x 16 B, then scales g x R.
template <int QT, int GS>
__global__ void pack_rk_kernel(const uint8_t* __restrict__ raw, const float* __restrict__ scales, uint8_t* __restrict__ packed,
                               const unsigned long long* __restrict__ cta_off, int M_total, int K, int row_base, int rows_src, int tt, int m, int cap) {
    using RK = Rk<QT, GS>;
    const int c = blockIdx.x, G = gridDim.x;
    const RkPart pt = rk_part(M_total, c, G, cap);
    const int kbytes = K * RK::ES, nkc = (kbytes + kStageRowBytes - 1) / kStageRowBytes, Gtot = K / GS;
    uint8_t* tile_base = packed + cta_off[c];
    for (int t = 0; t < pt.nt; ++t) {
        int lr0, R;
        rk_tile(pt, t, lr0, R);
        const int sb = RK::stage_bytes(R);
        for (int kc = blockIdx.y; kc < nkc; kc += gridDim.y) {
            uint8_t* st = tile_base + ((size_t)kc * tt + m) * sb;
            const int n_items = R * RK::PIECES + R * RK::GPS;
            for (int idx = threadIdx.x; idx < n_items; idx += blockDim.x) {
                if (idx < R * RK::PIECES) {
                    const int pc = idx / R, lane = idx - pc * R;
                    const int srow = pt.rb + lr0 + lane - row_base;
                    if (srow < 0 || srow >= rows_src) continue;
                    const int eb = kc * kStageRowBytes + pc * 16;
                    uint4 v = make_uint4(0, 0, 0, 0);
                    if (eb < kbytes) v = *reinterpret_cast<const uint4*>(raw + (size_t)srow * kbytes + eb);
                    *reinterpret_cast<uint4*>(st + (size_t)idx * 16) = v;
                } else {
                    const int si = idx - R * RK::PIECES, g = si / R, lane = si - g * R;
                    const int srow = pt.rb + lr0 + lane - row_base;
                    if (srow < 0 || srow >= rows_src) continue;
                    const int gi = kc * RK::GPS + g;
                    *reinterpret_cast<float*>(st + (size_t)R * kStageRowBytes + (size_t)si * 4) = gi < Gtot ? scales[(size_t)srow * Gtot + gi] : 0.0f;
                }
            }
        }
        tile_base += (size_t)nkc * tt * sb;
    }
}

// Phase p of a token (p = 4*layer + {0 QKV, 1 Wo, 2 W1/W3, 3 W2}, p = 4*n_layers: classifier): where its weights are and how
// they are cut.  Pure function of (params, p): the producer and every consumer warp evaluate it independently and agree.
struct PhaseShape {
    int tt;                            // sub-streams per tile (2 for the fused W1/W3 matrix)
    int K;                             // input length
    int M;                             // output rows
};

__device__ __forceinline__ PhaseShape phase_shape(const MegaParams& p, int pi) {
    PhaseShape s;
    if (pi == 4 * p.n_layers) { s.tt = 1; s.K = p.dim; s.M = p.vocab; return s; }
    const int ph = pi & 3;
    if (ph == 0)      { s.tt = 1; s.K = p.dim;    s.M = p.qkv_rows; }
    else if (ph == 1) { s.tt = 1; s.K = p.dim;    s.M = p.dim; }
    else if (ph == 2) { s.tt = 2; s.K = p.dim;    s.M = p.hidden; }
    else              { s.tt = 1; s.K = p.hidden; s.M = p.dim; }
    return s;
}
// start of this CTA's weight stream of phase pi: two dependent global loads (layer table, offset table).  Evaluated once per
// phase into a shared-memory table at kernel start - in the producer's loop they were a 2 us bubble in the stream per phase.
__device__ __forceinline__ const uint8_t* phase_stream(const MegaParams& p, int pi) {
    if (pi == 4 * p.n_layers) return p.cls + p.off_cls[blockIdx.x];
    const MegaLayer* L = p.layers + (pi >> 2);
    const int ph = pi & 3;
    if (ph == 0) return L->qkv + p.off_qkv[blockIdx.x];
    if (ph == 1) return L->wo + p.off_wo[blockIdx.x];
    if (ph == 2) return L->w13 + p.off_w13[blockIdx.x];
    return L->w2 + p.off_w2[blockIdx.x];
}

// Geometry of the five phase kinds (QKV, Wo, W1/W3, W2, classifier) for THIS CTA, computed once at kernel start into shared
// memory: the divisions behind rk_part / rk_tile / the superblock split cost ~1 us per phase when redone by every warp.
constexpr int kGeomStride = 32;     // ints per kind: M, K, tt, rb, nr, nt, nkc, sk, nsb, then nt + 1 tile boundaries (local rows)
constexpr int kGeomMaxTiles = 20;
enum { PG_M = 0, PG_K, PG_TT, PG_RB, PG_NR, PG_NT, PG_NKC, PG_SK, PG_NSB, PG_LR };
template <int QT, int GS>
__device__ __forceinline__ void fill_geometry(const MegaParams& p, int* geom, int kind) {
    using RK = Rk<QT, GS>;
    const PhaseShape s = phase_shape(p, kind == 4 ? 4 * p.n_layers : kind);
    const RkPart pt = rk_part(s.M, blockIdx.x, gridDim.x, p.tile_cap);
    int* g = geom + kind * kGeomStride;
    g[PG_M] = s.M; g[PG_K] = s.K; g[PG_TT] = s.tt; g[PG_RB] = pt.rb; g[PG_NR] = pt.nr; g[PG_NT] = pt.nt;
    g[PG_NKC] = ceil_div(s.K * RK::ES, kStageRowBytes);
    g[PG_SK] = kPairGroups / s.tt / RK::GPS;
    g[PG_NSB] = ceil_div(g[PG_NKC], g[PG_SK]);
    for (int t = 0; t <= pt.nt && t <= kGeomMaxTiles; ++t) g[PG_LR + t] = pt.nt ? pt.nr * t / pt.nt : 0;
}

// ---------------------------------------------------------------------------------------------- activation rebuild
__device__ __forceinline__ float group_max8(float m) {
    m = fmaxf(m, __shfl_xor_sync(kFull, m, 1));
    m = fmaxf(m, __shfl_xor_sync(kFull, m, 2));
    m = fmaxf(m, __shfl_xor_sync(kFull, m, 4));
    return m;
}

// quantise one lane's share of a group (PER consecutive values, 8 lanes per group) given the group's max |value|:
// quant::quantize (quant_operators.cpp:26-47); packed and stored as words.
// Which values of a 64-wide (32-wide) quantisation group one of its 8 lanes owns.  Interleaved (default): the lane's q-th pair
// of values is the pair 8 q + sub of the group, so the 8 lanes of a group read 128 contiguous bytes of tagged words per load
// and a warp's LDG.128 covers whole 32-byte sectors.  (Round 1 gave a lane 8 consecutive values = 64 contiguous bytes: its
// four loads each touched half a sector, i.e. every sector of an exchanged vector crossed the L2 -> SM fabric twice.)
#ifndef FL_NO_POLL_IL
constexpr bool kPollIL = true;
#else
constexpr bool kPollIL = false;
#endif
template <int GS>
__host__ __device__ __forceinline__ int lane_elem(int sub, int q) { return kPollIL ? 16 * q + 2 * sub : sub * (GS / 8) + 2 * q; }      // first of the pair, inside the group

template <int QT, int GS>
__device__ __forceinline__ void quant_store(uint8_t* xq, float* xs, const float (&y)[GS / 8], float m, int g, int sub, float* tap) {
    constexpr int PER = GS / 8;
    const float QF = (QT == Q_INT8) ? 127.0f : 5792.0f;
    const float sc = __fdiv_rn(m, QF);
    if (sub == 0) xs[g] = sc;
    uint32_t qv[PER];
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        // (T)(y / sc) with a true IEEE division, inlined.  Measured alternatives, both slower: a reciprocal-multiply fast path
        // with a fallback near integers (some lane nearly always needs the fallback), and a non-inlined helper per value / per
        // 4 values (the calls cost more than the instruction-cache space they save).
        qv[i] = (uint32_t)cvtt_x86(__fdiv_rn(y[i], sc)) & ((QT == Q_INT8) ? 0xffu : 0xffffu);
    }
    if constexpr (kPollIL) {
        // pairs of consecutive elements: 16-bit (INT8) / 32-bit (INT16) stores, natural element order
#pragma unroll
        for (int q = 0; q < PER / 2; ++q) {
            const int e = g * GS + lane_elem<GS>(sub, q);
            if (QT == Q_INT8) *reinterpret_cast<uint16_t*>(xq + e) = (uint16_t)(qv[2 * q] | (qv[2 * q + 1] << 8));
            else *reinterpret_cast<uint32_t*>(xq + (size_t)e * 2) = qv[2 * q] | (qv[2 * q + 1] << 16);
            if (tap) { tap[e] = y[2 * q]; tap[e + 1] = y[2 * q + 1]; }
        }
    } else {
        const int e0 = g * GS + sub * PER;
        constexpr int EPW = (QT == Q_INT8) ? 4 : 2;            // elements per 32-bit word
        uint32_t pk[PER / EPW];
#pragma unroll
        for (int i = 0; i < PER; ++i) pk[i / EPW] = (i % EPW == 0) ? qv[i] : (pk[i / EPW] | (qv[i] << ((32 / EPW) * (i % EPW))));
        uint32_t* dst = reinterpret_cast<uint32_t*>(xq + (size_t)e0 * ((QT == Q_INT8) ? 1 : 2));     // natural element order
#pragma unroll
        for (int i = 0; i < PER / EPW; ++i) dst[i] = pk[i];
        if (tap) {
#pragma unroll
            for (int i = 0; i < PER; ++i) tap[e0 + i] = y[i];
        }
    }
}

// simd::rmsnorm's sum of squares (x86_simd.cpp:941-962 via the __AVX2 typo at :1093): four FMA chains over x[4i+j], then
// 0 + l0 + l1 + l2 + l3.  xt is the TRANSPOSED fp32 vector in shared memory: xt[j * n/4 + i] = x[4i + j], so lane j
// streams its chain with 16-byte loads, double-buffered (profiles/r01: 2x the
// natural-layout version).  Called by one warp; returns the value in all its lanes.
__device__ __forceinline__ float sumsq_chain_t(const float* xt, int n, int lane) {
    float acc = 0.0f;
    if (lane < 4) {
        const int nv = n >> 4;                                // float4s per chain
        const float4* p = reinterpret_cast<const float4*>(xt + lane * (n >> 2));
        auto fma8 = [](const float4 (&v)[8], float acc) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                acc = __fmaf_rn(v[u].x, v[u].x, acc); acc = __fmaf_rn(v[u].y, v[u].y, acc);
                acc = __fmaf_rn(v[u].z, v[u].z, acc); acc = __fmaf_rn(v[u].w, v[u].w, acc);
            }
            return acc;
        };
        int i = 0;
        if (nv >= 8) {
            // two register sets in turn (no copies, no clamped indices): the loads of one unit of 8 float4s are in flight
            // while the other unit's 32 dependent FMAs run
            float4 a[8], b[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) a[u] = p[u];
#pragma unroll 1
            while (i + 16 <= nv) {
#pragma unroll
                for (int u = 0; u < 8; ++u) b[u] = p[i + 8 + u];
                acc = fma8(a, acc);
                if (i + 24 <= nv) {
#pragma unroll
                    for (int u = 0; u < 8; ++u) a[u] = p[i + 16 + u];
                }
                acc = fma8(b, acc);
                i += 16;
            }
            if (i + 8 <= nv) { acc = fma8(a, acc); i += 8; }
        }
#pragma unroll 1
        for (; i < nv; ++i) {
            const float4 v = p[i];
            acc = __fmaf_rn(v.x, v.x, acc); acc = __fmaf_rn(v.y, v.y, acc);
            acc = __fmaf_rn(v.z, v.z, acc); acc = __fmaf_rn(v.w, v.w, acc);
        }
    }
    const float l0 = __shfl_sync(kFull, acc, 0), l1 = __shfl_sync(kFull, acc, 1);
    const float l2 = __shfl_sync(kFull, acc, 2), l3 = __shfl_sync(kFull, acc, 3);
    float res = __fadd_rn(0.0f, l0);
    res = __fadd_rn(res, l1);
    res = __fadd_rn(res, l2);
    res = __fadd_rn(res, l3);
    return res;
}

// Rebuild the quantised activation vector of a phase in shared memory (every CTA, redundantly) from the tagged vector
// `src` (K = dim elements: the W2 input goes through build_hd below), waiting for every word to carry `tag`.
//   gain != NULL: y = (x*w)*r, r = 1/sqrt(mean(x^2)+eps) (simd::rmsnorm, x86_simd.cpp:1754);  gain == NULL: y = x.
// Thread layout: 8 lanes per quantisation group, 32 groups per pass, at most MAXP passes (K <= 24 * 256 = 6144, checked by
// the host); the values wait in registers.
// Polling.  A CTA that finishes its drain early would otherwise re-read the whole vector (32 KB of tagged words, 12 loads of
// 16 bytes per thread) once per L2 round trip until the slowest producer has published - several rounds on 148 CTAs, i.e. more
// L2 traffic than the weights the slow CTAs are still streaming.  So a thread first polls ONE word per pass, the SENTINEL: the
// last row of the producer CTA that owns its last word (`n_prod` producers own dim * c / n_prod ... rows each and publish
// a tile's rows with one store instruction).  The lanes of a warp share ~9 sentinels, which coalesce into as many 32-byte
// sectors: a waiting round costs ~7 KB per CTA instead of 48 KB.  Only when its sentinel carries the tag does the thread
// load its own words, which are then valid at the first look except at producer boundaries (re-polled as before).  The
// sentinel's index only has to be SOME later word of the same vector, so it is computed in floating point.
// RX (FL_FLAG_RELAXED, measurement of what bit-exactness costs): the sum of squares is a tree reduction instead of the
// reference's four 1024-step FMA chains - same value up to FP32 rounding (~1e-7 relative), NOT the reference's bits.
// (Round 2 also measured the chain on the otherwise idle chain warp, following the poll pass by pass: 386 against 438 tokens/s,
// profiles/r02/ab_chain_warp.log; the code is in the history, commit 777712b.)
template <int QT, int GS, bool RX = false>
__device__ __forceinline__ void build_activation(uint8_t* xq, float* xs, float* xt, float* misc, const uint2* src, uint32_t tag,
                                                 const float* gain, int K, int n_prod, float* tap, int tid, Prof& pf, uint32_t* gate, uint32_t gate_val,
                                                 uint32_t chain_phases) {
    using RK = Rk<QT, GS>;
    constexpr int PER = GS / 8;                 // values per thread per group
    constexpr int LPP = PER / 2;                // 16-byte loads per thread per pass
    constexpr int GPP = kConsumerThreads / 8;   // groups per pass
    constexpr int MAXP = 24 / PER;              // passes (24 values per thread: registers, not shared memory)
    const int warp = tid >> 5, lane = tid & 31;
    const int kpad_bytes = ceil_div(K * RK::ES, kStageRowBytes) * kStageRowBytes;      // the image is padded to whole stages
    const int G = K / GS;
    const int sub = tid & 7, g0 = tid >> 3;
    const int n_pass = ceil_div(G, GPP);
    uint4 w[MAXP][LPP];
    uint2 sv[MAXP];                             // the sentinel of each pass
    const float rows_per_prod = (float)K / (float)n_prod, prod_per_row = (float)n_prod / (float)K;
    auto sentinel = [&](int ps) {
        const int last = (g0 + ps * GPP) * GS + lane_elem<GS>(sub, LPP - 1) + 1;        // my last word of the pass
        const int c = (int)((float)(last + 1) * prod_per_row);                           // ~ its producer
        return max(last, min(K - 1, (int)((float)(c + 1) * rows_per_prod) - 1));         // ~ that producer's last row
    };
    // Sentinels pay where the wait is long (x1 after Wo / W2: every CTA waits for the slowest drain); where the words are
    // mostly there already (the attention output, the quantised hd) the extra round trip costs more than the saved traffic
    // (measured with sentinels everywhere: 450 against 461 tokens/s, profiles/r02/ab_sentinel_all.log)
    const bool use_sentinel = gain != nullptr;
    uint32_t loaded = use_sentinel ? 0u : 0xffffffffu;      // bit ps: the pass's own words have been requested
#pragma unroll
    for (int ps = 0; ps < MAXP; ++ps) {
        const int g = g0 + ps * GPP;
        if (ps < n_pass && g < G) {
            if (use_sentinel) sv[ps] = ld_tag1(src + sentinel(ps));
            else {
                const uint2* s = src + g * GS;
#pragma unroll
                for (int q = 0; q < LPP; ++q) w[ps][q] = ld_tag2(s + lane_elem<GS>(sub, q));
            }
        }
    }
    float2 gw[MAXP][LPP];
    auto load_gain = [&]() {
        // the gain vector (16 KB per layer, L2-resident)
#pragma unroll
        for (int ps = 0; ps < MAXP; ++ps) {
            const int g = min(g0 + ps * GPP, G - 1);        // clamped: always a valid address, unused where the thread has no group
#pragma unroll
            for (int q = 0; q < LPP; ++q) gw[ps][q] = __ldg(reinterpret_cast<const float2*>(gain + g * GS + lane_elem<GS>(sub, q)));
        }
    };
#if !defined(FL_OLD_REBUILD) && !defined(FL_GAIN_EARLY)
    // rmsnorm rebuilds with the chain on the chain warp: x * w runs under the chain, so the gain is fetched only after the poll
    // (its round trip hides under the chain as well, and its 24 registers are not held - and spilled - across the poll)
    constexpr bool kGainLate = !RX && kRbMode == 0;
#else
    constexpr bool kGainLate = false;
#endif
    // otherwise before the poll; measured in round 1: fetching it after the poll puts the products behind the chain and costs
    // more than the spill of half of it does
    if (gain && !kGainLate) load_gain();
    bool again;
    do {
        again = false;
#pragma unroll
        for (int ps = 0; ps < MAXP; ++ps) {
            const int g = g0 + ps * GPP;
            if (ps < n_pass && g < G) {
                const uint2* s = src + g * GS;
                if (!((loaded >> ps) & 1u)) {
                    if (sv[ps].y == tag) {
#pragma unroll
                        for (int q = 0; q < LPP; ++q) w[ps][q] = ld_tag2(s + lane_elem<GS>(sub, q));
                        loaded |= 1u << ps;
                    } else {
                        sv[ps] = ld_tag1(src + sentinel(ps));
                    }
                    again = true;               // the words just requested are looked at in the next round
                } else {
#pragma unroll
                    for (int q = 0; q < LPP; ++q)
                        if (w[ps][q].y != tag || w[ps][q].w != tag) { w[ps][q] = ld_tag2(s + lane_elem<GS>(sub, q)); again = true; }
                }
            }
        }
        // a poll that failed is not worth repeating at once: the LSU is shared with warps still working (measured again in
        // round 2: no sleep 427.8 against 431.0 tokens/s)
#ifndef FL_ISSUED_SLEEP
#define FL_ISSUED_SLEEP 40
#endif
#ifndef FL_GATE_SLEEP
#define FL_GATE_SLEEP 64
#endif
#ifndef FL_SENT_SLEEP
#define FL_SENT_SLEEP 30       // a sentinel round is 9 sectors per warp: a short back-off is enough (30 vs 100 ns: 492.9 vs 490.9 tokens/s, profiles/r02/ab_sentinel_sleep.log)
#endif
        if (again) __nanosleep(use_sentinel ? FL_SENT_SLEEP : 100);
    } while (again);
    pf.stop(tid, 0);
    pf.log(tid & 31, tid >> 5, 9, 0);
    pf.mark(tid, pf.trace_slot, pf.trace_slot >= 0);
#ifndef FL_OLD_REBUILD
    // rmsnorm rebuilds: the sum-of-squares chain runs on the CTA's chain warp (idle between two drains), and everything of
    // the rebuild that does not need its result (x * w, the group maxima) runs under it.  A consumer thread that has its words
    // stores the raw values into the transposed vector and ARRIVES on a named barrier (it does not wait); the chain warp waits
    // there for all 256, opens the producer's gate, walks the chain and arrives on a second barrier the consumers wait on before
    // the quantisation tail.  (Round 1/2a had the chain on consumer warp 7 between two consumer-wide barriers, with x * w and
    // the maxima in front of it: 1.2 us more per rebuild on the critical path, profiles/r02/trace_rebuild_*.log.)
    const bool chained = gain != nullptr && !RX;
#else
    const bool chained = false;
#endif
    if (chained) {
        // xt lies on the pair buffers: this CTA's own chain warp must have finished the previous drain (it nearly always has:
        // its rows are part of the vector just polled)
        while ((int)(ld_shared_volatile_u32(reinterpret_cast<const uint32_t*>(misc) + 29) - chain_phases) < 0) __nanosleep(20);
#ifdef FL_GATE_EARLY      // A/B: the first warp that has its words opens the gate (the others' last polls then share the memory system with the refill)
        if (lane == 0 && gate) st_shared_volatile_u32(gate, gate_val);
        if (tid == 0) st_shared_volatile_u32(reinterpret_cast<uint32_t*>(misc) + 19, 0u);
#else
        if (tid == 0) st_shared_volatile_u32(reinterpret_cast<uint32_t*>(misc) + 19, gate ? gate_val : 0u);       // the chain warp opens the gate
#endif
    } else {
        // every warp has left the previous phase (its drain / attention read the image this build overwrites)
        consumer_sync();
        pf.log(lane, warp, 20, 0);
        // the input is here: the producer may prefetch again (measured: releasing later costs more ring prefetch than it saves in contention)
        if (gate && tid == 0) st_shared_volatile_u32(gate, gate_val);
        // zero the padded tail so padded groups contribute fma(0, 0, acc) == acc
        for (int i = K * RK::ES + tid * 4; i < kpad_bytes; i += kConsumerThreads * 4) *reinterpret_cast<uint32_t*>(xq + i) = 0u;
        for (int i = G + tid; i < (kpad_bytes / kStageRowBytes) * RK::GPS; i += kConsumerThreads) xs[i] = 0.0f;
    }
    // values out of the tagged words (the tags' registers are free from here on)
    float y[MAXP][PER];
    float ss_part = 0.0f;
#pragma unroll
    for (int ps = 0; ps < MAXP; ++ps) {
        const int g = g0 + ps * GPP;
        if (ps < n_pass && g < G) {
#pragma unroll
            for (int q = 0; q < LPP; ++q) { y[ps][2 * q] = __uint_as_float(w[ps][q].x); y[ps][2 * q + 1] = __uint_as_float(w[ps][q].z); }
            if (gain && RX) {
#pragma unroll
                for (int i = 0; i < PER; ++i) ss_part = __fmaf_rn(y[ps][i], y[ps][i], ss_part);
            }
            if (gain && !RX) {
                // raw x -> transposed vector for the chain (xt[j * K/4 + i] = x[4 i + j])
                if constexpr (kPollIL) {
#pragma unroll
                    for (int q = 0; q < LPP; ++q) {
                        const int e = g * GS + lane_elem<GS>(sub, q);           // even
                        float* d = xt + (e & 3) * (K >> 2) + (e >> 2);
                        d[0] = y[ps][2 * q]; d[K >> 2] = y[ps][2 * q + 1];
                    }
                } else {
                    const int e0 = g * GS + sub * PER;                   // multiple of 4
                    if constexpr (PER == 8) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) *reinterpret_cast<float2*>(xt + j * (K >> 2) + (e0 >> 2)) = make_float2(y[ps][j], y[ps][4 + j]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) xt[j * (K >> 2) + (e0 >> 2)] = y[ps][j];
                    }
                }
            }
        }
    }
    // x*w and the group maxima (max |(x*w)*r| == (max |x*w|)*r: rounding is monotonic, r > 0)
    float m[MAXP];
    auto pre = [&]() {
#pragma unroll
        for (int ps = 0; ps < MAXP; ++ps) {
            const int g = g0 + ps * GPP;
            float mm = 0.0f;
            if (ps < n_pass && g < G) {
                if (gain) {
#pragma unroll
                    for (int q = 0; q < LPP; ++q) {                     // x*w, multiply_avx256 x86_simd.cpp:1359
                        y[ps][2 * q] = __fmul_rn(y[ps][2 * q], gw[ps][q].x); y[ps][2 * q + 1] = __fmul_rn(y[ps][2 * q + 1], gw[ps][q].y);
                    }
                }
#pragma unroll
                for (int i = 0; i < PER; ++i) mm = fmaxf(mm, fabsf(y[ps][i]));
            }
            m[ps] = group_max8(mm);
        }
    };
    bool have_rr = false;
    if (chained) {
        delay_point<5>();
        if (kRbMode == 0) {
            asm volatile("bar.arrive 5, %0;" :: "n"(kConsumerThreads + 32) : "memory");       // my part of xt is written
            pf.log(lane, warp, 20, 0);
            if (kGainLate) load_gain();
            pre();
        } else if (kRbMode == 1) {
            pre();
            asm volatile("bar.arrive 5, %0;" :: "n"(kConsumerThreads + 32) : "memory");
            pf.log(lane, warp, 20, 0);
        } else if (warp == kSerialWarp) {
            // mode 2: the chain stays on consumer warp 7 (alone on its sub-partition but for warp 3); its own x * w follows the chain
            asm volatile("bar.sync 5, %0;" :: "n"(kConsumerThreads) : "memory");
            if (lane == 0 && gate) st_shared_volatile_u32(gate, gate_val);
            const float ss = sumsq_chain_t(xt, K, lane);
            if (lane == 0) misc[18] = rms_scale(ss, K);
            __syncwarp();
            asm volatile("bar.arrive 6, %0;" :: "n"(kConsumerThreads) : "memory");
            pre();
            have_rr = true;
        } else {
            asm volatile("bar.arrive 5, %0;" :: "n"(kConsumerThreads) : "memory");
            pre();
        }
    } else {
        pre();
    }
    pf.stop(tid, 8);
    pf.log(lane, warp, 21, 0);
    float rr = 1.0f;
    if (gain && RX) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ss_part = __fadd_rn(ss_part, __shfl_xor_sync(kFull, ss_part, o));
        if (lane == 0) misc[8 + warp] = ss_part;
        consumer_sync();
        float ss = misc[8];
#pragma unroll
        for (int w8 = 1; w8 < kConsumerWarps; ++w8) ss = __fadd_rn(ss, misc[8 + w8]);
        rr = rms_scale(ss, K);
    } else if (chained) {
        if (kRbMode != 2) asm volatile("bar.sync 6, %0;" :: "n"(kConsumerThreads + 32) : "memory");         // the chain warp has published the scale
        else if (!have_rr) asm volatile("bar.sync 6, %0;" :: "n"(kConsumerThreads) : "memory");
        rr = misc[18];
        // every consumer warp is here, i.e. has left the previous phase: the image may be overwritten
        for (int i = K * RK::ES + tid * 4; i < kpad_bytes; i += kConsumerThreads * 4) *reinterpret_cast<uint32_t*>(xq + i) = 0u;
        for (int i = G + tid; i < (kpad_bytes / kStageRowBytes) * RK::GPS; i += kConsumerThreads) xs[i] = 0.0f;
    } else if (gain) {
        consumer_sync();
        pf.log(lane, warp, 22, 0);
        if (warp == kSerialWarp) {
            const float ss = sumsq_chain_t(xt, K, lane);
            pf.log(lane, warp, 23, 0);
            if (lane == 0) misc[18] = rms_scale(ss, K);
        }
        consumer_sync();
        rr = misc[18];
    }
    pf.stop(tid, 9);
    pf.log(lane, warp, 24, 0);
#pragma unroll
    for (int ps = 0; ps < MAXP; ++ps) {
        const int g = g0 + ps * GPP;
        if (ps < n_pass && g < G) {
            if (gain) {
#pragma unroll
                for (int i = 0; i < PER; ++i) y[ps][i] = __fmul_rn(y[ps][i], rr);     // (x*w)*r
                m[ps] = __fmul_rn(m[ps], rr);
            }
            quant_store<QT, GS>(xq, xs, y[ps], m[ps], g, sub, tap);
        }
    }
    pf.log(lane, warp, 25, 0);
    if (chained) delay_point<1>();
    consumer_sync();
}

// The W2 input (hd, K = hidden) is the one exchange whose size hurts: as tagged fp32 words it is 8 bytes per element read by
// every CTA (88 KB per CTA, 13 MB through the L2 per layer at the 7B shape, ~2 us of L2 bandwidth alone) followed by the longest
// quantisation tail.  It needs no vector-wide quantity (no rmsnorm), so it is quantised ONCE: CTA c is the designated quantiser
// of groups [G c / n, G (c + 1) / n) - it polls their raw words (published by the one or two CTAs that own those W1/W3 rows),
// quantises them exactly as quant::quantize does (quant_operators.cpp:26-47) and publishes payload words and scale as tagged
// words; every CTA then polls the quantised vector (K * ES / 4 payload words, then G scales: 23 KB instead of 88 KB) straight
// into its shared-memory image.  One more L2 hop, a quarter of the bytes and the whole quantisation tail less.
template <int QT, int GS>
__device__ __forceinline__ void build_hd(uint8_t* xq, float* xs, const uint2* raw, uint2* qt, uint32_t tag_raw, uint32_t tag_q, int K,
                                         int tid, Prof& pf, uint32_t* gate, uint32_t gate_val) {
    using RK = Rk<QT, GS>;
    constexpr int PER = GS / 8;                             // values per thread of a quantiser group (8 lanes per group)
    constexpr int EPW = (QT == Q_INT8) ? 4 : 2;             // elements per payload word
    constexpr int PW = GS / EPW;                            // payload words per group
    const int G = K / GS;
    const int NPW = G * PW;                                 // payload words of the vector (even)
    const int sub = tid & 7;
    // ---- A: quantise my groups
    const int g_lo = (int)((long long)G * blockIdx.x / gridDim.x), g_hi = (int)((long long)G * (blockIdx.x + 1) / gridDim.x);
    const unsigned gmask = 0xffu << (tid & 24);             // the 8 lanes of a group stay together
#pragma unroll 1
    for (int g = g_lo + (tid >> 3); g < g_hi; g += kConsumerThreads / 8) {
        const uint2* s = raw + g * GS + sub * PER;
        uint4 w[PER / 2];
#pragma unroll
        for (int q = 0; q < PER / 2; ++q) w[q] = ld_tag2(s + 2 * q);
        bool again;
        do {
            again = false;
#pragma unroll
            for (int q = 0; q < PER / 2; ++q)
                if (w[q].y != tag_raw || w[q].w != tag_raw) { w[q] = ld_tag2(s + 2 * q); again = true; }
        } while (__any_sync(gmask, again));
        float y[PER];
        float m = 0.0f;
#pragma unroll
        for (int q = 0; q < PER / 2; ++q) { y[2 * q] = __uint_as_float(w[q].x); y[2 * q + 1] = __uint_as_float(w[q].z); }
#pragma unroll
        for (int i = 0; i < PER; ++i) m = fmaxf(m, fabsf(y[i]));
        m = fmaxf(m, __shfl_xor_sync(gmask, m, 1));
        m = fmaxf(m, __shfl_xor_sync(gmask, m, 2));
        m = fmaxf(m, __shfl_xor_sync(gmask, m, 4));
        const float sc = __fdiv_rn(m, (QT == Q_INT8) ? 127.0f : 5792.0f);
        if (sub == 0) st_tag(qt + NPW + g, sc, tag_q);
        uint32_t pk[PER / EPW];
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            const uint32_t q = (uint32_t)cvtt_x86(__fdiv_rn(y[i], sc)) & ((QT == Q_INT8) ? 0xffu : 0xffffu);
            pk[i / EPW] = (i % EPW == 0) ? q : (pk[i / EPW] | (q << ((32 / EPW) * (i % EPW))));
        }
#pragma unroll
        for (int i = 0; i < PER / EPW; ++i) st_tag(qt + (size_t)g * PW + sub * (PER / EPW) + i, __uint_as_float(pk[i]), tag_q);
    }
    // this CTA's share of the quantisation is out (the one poll others wait for): the producer may refill the ring during the
    // second hop - the sync also says that every warp has left the previous phase (measured: 447.6 against 431.0 tokens/s with
    // the release after the second poll, profiles/r02/ab_hd_gate.log)
    consumer_sync();
    if (gate && tid == 0) st_shared_volatile_u32(gate, gate_val);
    {   // zero the padded tail of the image so padded groups contribute fma(0, 0, acc) == acc
        const int kpad_bytes = ceil_div(K * RK::ES, kStageRowBytes) * kStageRowBytes;
        for (int i = K * RK::ES + tid * 4; i < kpad_bytes; i += kConsumerThreads * 4) *reinterpret_cast<uint32_t*>(xq + i) = 0u;
        for (int i = G + tid; i < (kpad_bytes / kStageRowBytes) * RK::GPS; i += kConsumerThreads) xs[i] = 0.0f;
    }
    // ---- B: the quantised vector -> shared-memory image.  Word pair i of [payload | scales] lands at word pair i of
    //      [xq | xs] (xs directly behind the payload in index space only: two arrays, one branch).
    const int NP = (NPW + G + 1) >> 1;                      // 16-byte pairs ([NPW payload words][G scale words], the buffer is padded to a pair)
    constexpr int MAXL = 6;                                 // loads per thread per batch (7B: 1462 pairs = 5.7 per thread)
    uint32_t* xq32 = reinterpret_cast<uint32_t*>(xq);
#pragma unroll 1
    for (int base = 0; base < NP; base += MAXL * kConsumerThreads) {
        uint4 w[MAXL];
#pragma unroll
        for (int l = 0; l < MAXL; ++l) {
            const int pi = base + l * kConsumerThreads + tid;
            if (pi < NP) w[l] = ld_tag2(qt + 2 * pi);
        }
        bool again;
        do {
            again = false;
#pragma unroll
            for (int l = 0; l < MAXL; ++l) {
                const int pi = base + l * kConsumerThreads + tid;
                if (pi < NP && (w[l].y != tag_q || (2 * pi + 1 < NPW + G && w[l].w != tag_q))) { w[l] = ld_tag2(qt + 2 * pi); again = true; }
            }
            if (again) __nanosleep(100);
        } while (again);
        pf.stop(tid, 0);
        pf.mark(tid, pf.trace_slot, pf.trace_slot >= 0 && base + MAXL * kConsumerThreads >= NP);
#pragma unroll
        for (int l = 0; l < MAXL; ++l) {
            const int wi = 2 * (base + l * kConsumerThreads + tid);
            if (wi < NPW) { xq32[wi] = w[l].x; xq32[wi + 1] = w[l].z; }      // NPW is even: a pair never straddles the two arrays
            else if (wi < NPW + G) { xs[wi - NPW] = __uint_as_float(w[l].x); if (wi + 1 < NPW + G) xs[wi + 1 - NPW] = __uint_as_float(w[l].z); }
        }
    }
    consumer_sync();
}

// ---------------------------------------------------------------------------------------------- consumer GEMV phase
// One stage = 256 bytes of each of the R rows of a tile.  Lane i owns row i: integer dots of its GPS groups (exact, any
// order) and the scale products ws * xs, stored as (product, float(dot)) pairs for the chain warp:
// acc = fma(ws * xs, float(dot), acc)  (quant_operators.cpp:274-275).  pairs: [group][lane] float2 of this stage's groups.
// xp / xsp: the activation image and scales of this K chunk (same for all lanes: broadcast loads).
template <int QT, int GS>
__device__ __forceinline__ void stage_pairs(const uint8_t* sp, int R, const uint4* xp, const float* xsp, int lane, float2* pairs) {
    using RK = Rk<QT, GS>;
    const uint4* wp = reinterpret_cast<const uint4*>(sp) + lane;
    const float* ssp = reinterpret_cast<const float*>(sp + R * kStageRowBytes) + lane;
#pragma unroll
    for (int g = 0; g < RK::GPS; ++g) {
        int dj[RK::PPG];
#pragma unroll
        for (int j = 0; j < RK::PPG; ++j) dj[j] = dot16<QT>(wp[(g * RK::PPG + j) * R], xp[g * RK::PPG + j], 0);
        int d = dj[0];
#pragma unroll
        for (int j = 1; j < RK::PPG; ++j) d += dj[j];
        pairs[g * 32 + lane] = make_float2(__fmul_rn(ssp[g * R], xsp[g]), __int2float_rn(d));
    }
}

// ---------------------------------------------------------------------------------------------- attention part
// execute_attn (transformer.cpp:397-455) for query head qh, CTA `part` of `cph`: scores for a contiguous share of
// the keys, score exchange through L2 (tagged words), full softmax (redundantly per part), PV chains for HS/cph head dims.
//
// Everything that does not depend on the new token is requested BEFORE the q/k/v rows are polled: the part's K rows
// (registers), its V column block (TMA bulk copies into a ring of 32-row chunks, the V cache is stored in column blocks of
// HS/cph for exactly this; up to ~10 chunks = 320 rows are in flight at once) and the RoPE table row.  The two serial
// sections - softmax's sum and the PV chains - run as register-staged FP32 chains.
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// The K sweep of a long context (attention_part, `big`): the part's cached K rows [tb, kc_end) arrive through the weight ring
// in chunks of 32 rows (issue_k in attention_part filled the first NKC slots); warp w takes rows 4w .. 4w + 3 of every chunk,
// 8 lanes per row as in the register path, and warp 0 refills the slot of chunk c - 1 once every warp has reported it done.
struct KSweepOut { uint32_t ph; float mloc; };
template <int HS>
__device__ __forceinline__ KSweepOut k_sweep_ring(float* kring, uint64_t* vfull, uint32_t* kprog, const float* kc, const float* q_s, const float* k_s,
                                               float* att, uint2* att_g, int tb, int kc_end, int te, int pos, int NKC, int nkch, int cph,
                                               float attn_scale, uint32_t tag_score, uint32_t ph, int tid) {
    constexpr int EPL = HS / 8, KR = 32, KROW = HS;       // unpadded rows: a chunk of 32 rows is ONE 16 KB bulk copy (4-way bank conflicts on the reads are cheaper than 32 copies)
    const int warp = tid >> 5, lane = tid & 31, rr = lane >> 3, j = lane & 7;
    float mloc = -INFINITY;
    const float* qj = q_s + j;
    auto score = [&](const float4 (&k4)[EPL / 4], int t, bool live) {
        float acc = 0.0f;
#pragma unroll
        for (int q = 0; q < EPL / 4; ++q) {
            acc = __fmaf_rn(k4[q].x, qj[8 * (4 * q)], acc);
            acc = __fmaf_rn(k4[q].y, qj[8 * (4 * q + 1)], acc);
            acc = __fmaf_rn(k4[q].z, qj[8 * (4 * q + 2)], acc);
            acc = __fmaf_rn(k4[q].w, qj[8 * (4 * q + 3)], acc);
        }
        float tot = 0.0f;
#pragma unroll
        for (int k = 0; k < 8; ++k) tot = __fadd_rn(tot, __shfl_sync(kFull, acc, (rr << 3) + k));
        if (j == 0 && live) {
            const float sc = __fmul_rn(tot, attn_scale);                // att.multiply(attn_scale), transformer.cpp:443
            att[t] = sc;
            mloc = fmaxf(mloc, sc);
            if (cph > 1) st_tag(att_g + t, sc, tag_score);
        }
    };
    auto issue_k = [&](int c) {                             // warp 0, all lanes (same as attention_part's)
        const int slot = c % NKC, t0 = tb + c * KR, rows = min(KR, kc_end - t0);
        if (lane == 0) {
            mbar_arrive_expect_tx(&vfull[slot], (uint32_t)(rows * HS * 4));
            bulk_g2s(kring + (size_t)slot * KR * KROW, kc + (size_t)t0 * HS, (uint32_t)(rows * HS * 4), &vfull[slot]);
        }
        __syncwarp();
    };
#pragma unroll 1
    for (int c = 0; c < nkch; ++c) {
        const int slot = c % NKC;
        mbar_wait(&vfull[slot], (ph >> slot) & 1u);
        ph ^= 1u << slot;
        const int t = tb + c * KR + warp * 4 + rr;
        const float4* kp = reinterpret_cast<const float4*>(kring + ((size_t)slot * KR + warp * 4 + rr) * KROW) + j;      // k_cache_index: conflict-free
        float4 k4[EPL / 4];
#pragma unroll
        for (int q = 0; q < EPL / 4; ++q) k4[q] = (t < kc_end) ? kp[8 * q] : make_float4(0.f, 0.f, 0.f, 0.f);
        score(k4, t, t < kc_end);
        __syncwarp();
        if (lane == 0) st_shared_volatile_u32(kprog + warp, (uint32_t)c + 1u);
        if (warp == 0 && c >= 1 && c - 1 + NKC < nkch) {
            // every warp has finished chunk c - 1: refill its slot
            while (!__all_sync(kFull, lane >= kConsumerWarps || (int)ld_shared_volatile_u32(kprog + lane) >= c)) { }
            issue_k(c - 1 + NKC);
        }
    }
    if (pos >= tb && pos < te && warp == kConsumerWarps - 1) {
        // the new token's key, from the RoPE'd row in shared memory
        float4 k4[EPL / 4];
#pragma unroll
        for (int q = 0; q < EPL / 4; ++q)
            k4[q] = make_float4(k_s[8 * (4 * q) + j], k_s[8 * (4 * q + 1) + j], k_s[8 * (4 * q + 2) + j], k_s[8 * (4 * q + 3) + j]);
        score(k4, pos, rr == 0);
    }
    return KSweepOut{ph, mloc};
}

template <int HS, bool RX = false, bool LC = false>
__device__ __forceinline__ void attention_part(const MegaParams& p, const SeqView sv, uint8_t* smem, int layer, int qh, int part, int pos, int bs,
                                               uint32_t tag_qkv, uint32_t tag_score, uint32_t tag_out, uint32_t phases_drained, int tid, Prof& pf,
                                               uint32_t* gate = nullptr, uint32_t gate_val = 0u) {
    constexpr int EPL = HS / 8;
    const int cph = p.cph;
    const int DW = HS / cph;                                // head dims owned by this part
    float* att = reinterpret_cast<float*>(smem + p.off_att);
    float* q_s = reinterpret_cast<float*>(smem + p.off_misc) + 32;
    float* k_s = q_s + HS;
    float* v_s = k_s + HS;
    float* red = reinterpret_cast<float*>(smem + p.off_misc);
    uint32_t* vphase = reinterpret_cast<uint32_t*>(smem + p.off_misc) + 28;      // bit s: parity of the next completion of V barrier s
    uint64_t* vfull = reinterpret_cast<uint64_t*>(smem + p.off_vbars);
    constexpr int VR = kVChunkRows;

    const int hgs = p.n_heads / p.n_kv_heads;
    const int kvh = qh / hgs, g = qh % hgs;
    const int dim = p.n_heads * HS, kv_dim = p.n_kv_heads * HS;
    const int n = pos + 1;
    const int warp = tid >> 5, lane = tid & 31;
    const size_t cache_off = ((size_t)layer * p.n_kv_heads + kvh) * p.max_seq * HS;
    float* kc = sv.kc + cache_off;
    float* vc = sv.vc + cache_off + (size_t)part * p.max_seq * DW;      // this part's column block: [max_seq / 4][DW][4 positions]
    const int d0 = part * DW;

    // ---- V ring: chunk c = cached rows [c*VR, min(pos, (c+1)*VR)) of the column block, one bulk copy each
    const int n_chunks = ceil_div(pos, VR);                 // the new row (t == pos) comes from v_s
    // Staging area.  Short contexts: the activation image + pair buffers (idle during attention), ~10 chunks, and the producer
    // refills the weight ring with Wo / W13 while softmax and PV run.  LONG contexts (more than twice that): the weight ring
    // itself - it is empty between the QKV drain and the moment this function releases the producer's gate - so up to 32 chunks
    // (~170 KB) of V are in flight and the sweep runs at the SM's ingest rate instead of at 43 KB per round trip; the gate then
    // opens after PV (Wo's 2.4 us of prefetch are noise next to a 10-30 us sweep).  Needs the gate: not in multi-sequence launches.
    // LC: the long-context code exists only in the kernel variant the host launches when a sequence can get that far (its
    // registers and instructions would otherwise tax the short-context path: +430 bytes of spills in the phase loop)
    // (the host launches the LC variant only for a single sequence that is already past 2 * n_vchunks chunks: launch_mega)
    constexpr bool big = LC;
#ifdef FL_LONG_K
    // A/B build: K through the ring as well - chunks of 32 rows = ONE 16 KB bulk copy each (the cache keeps a head's rows
    // contiguous), ~10 chunks in flight
    constexpr bool kRingK = true;
#else
    // Measured at the 13B shape from context 2048 / 7B from 900 (profiles/r02/longctx_k_ring.log, longctx_k_layout.log):
    //   register-staged K, lane j's 16 values contiguous in the row (round 1 layout) ............ 152.0 / 404.6 tokens/s
    //   K through the ring, one 512-byte copy per row into padded rows (32-lane ELECT waterfall) . 107   (longctx_ab.log)
    //   K through the ring, one 16 KB copy per chunk, unpadded rows (4-way bank conflicts) ....... 161.4 / 411.1
    //   new row layout (k_cache_index: a row's 8 lanes read 128 contiguous bytes), ring .......... 169.9 / 422.2
    //   new row layout, register-staged K ....................................................... 172.7 / 431.0   <- default
    // The register sweep had been slow because each LDG.128 of a warp touched 32 separate 16-byte pieces 64 bytes apart.
    constexpr bool kRingK = false;
#endif
    constexpr bool bigk = big && kRingK;
    float* v_stage = reinterpret_cast<float*>(smem + (big ? p.off_ring : p.off_vstage));
    const int NCH = big ? min(32, (p.n_slots * p.slot_bytes) / (VR * DW * 4)) : p.n_vchunks;
    auto issue_v = [&](int c) {                             // one thread
        const uint32_t slot = (uint32_t)c % (uint32_t)NCH;
        const uint32_t bytes = (uint32_t)((min(VR, pos - c * VR) + 3) >> 2) * DW * 16;      // whole blocks of 4 positions
        mbar_arrive_expect_tx(&vfull[slot], bytes);
        bulk_g2s(v_stage + (size_t)slot * VR * DW, vc + (size_t)c * VR * DW, bytes, &vfull[slot]);
    };
    // K/V rows of earlier tokens were appended by plain stores of another CTA (part 0 of this head); that CTA's chain warp
    // fenced them before it published tagged rows this CTA has polled since (see the Wo epilogue).  Acquire side of that
    // hand-off, before the cache is read through ld.cg (K) and through the async proxy (V bulk copies):
#ifndef FL_NO_KV_FENCE      // A/B build only: what the KV hand-off fences cost
    __threadfence();
#endif
    const int pvt = tid - (kConsumerThreads - DW);          // index inside the PV group (the last DW consumer threads), < 0 for the others
    uint32_t ph = *vphase;                                  // every thread follows the barriers' phases (same waits in the same order)
    if (pvt == 0 && !bigk) {
        // the ring also covers the pair buffers: wait until this CTA's chain warp has finished the QKV phase (it trails the
        // consumers by one superblock at most)
        while ((int)(ld_shared_volatile_u32(reinterpret_cast<uint32_t*>(smem + p.off_misc) + 29) - phases_drained) < 0) __nanosleep(50);
        fence_proxy_async();                                // the ring aliases memory the generic proxy wrote (activation image, pairs)
        asm volatile("fence.proxy.async.global;" ::: "memory");      // the V rows were written through the generic proxy
        for (int c = 0; c < min(NCH, n_chunks); ++c) issue_v(c);
    }

    // ---- this part's keys; the first batches of K rows are requested before the q/k/v rows are polled
    const int per = ceil_div(ceil_div(n, cph), 4) * 4;
    const int tb = part * per, te = min(n, tb + per);
    const int rr = lane >> 3, j = lane & 7;
    // long contexts: the part's cached K rows [tb, min(te, pos)) stream through the weight ring as well (before V, which is
    // requested once the sweep is over): chunks of 32 contiguous rows, one bulk copy each (the unpadded rows cost 4-way bank
    // conflicts on the LDS.128 of 4 keys x 8 lanes - cheaper than 32 copies), NKC chunks in flight instead of one register
    // batch per round trip
    constexpr int KR = 32, KROW = HS;                       // rows per K chunk; floats per staged row (unpadded: one bulk copy per chunk)
    const int kc_end = min(te, pos);
    const int nkch = bigk ? ceil_div(max(kc_end - tb, 0), KR) : 0;
    const int NKC = bigk ? min(32, (p.n_slots * p.slot_bytes) / (KR * KROW * 4)) : 1;
    float* kring = reinterpret_cast<float*>(smem + p.off_ring);
    uint32_t* kprog = reinterpret_cast<uint32_t*>(smem + p.off_misc) + 416;     // [warp] K chunks this warp has finished
    auto issue_k = [&](int c) {                             // warp 0, all lanes
        const int slot = c % NKC, t0 = tb + c * KR, rows = min(KR, kc_end - t0);
        if (lane == 0) {
            mbar_arrive_expect_tx(&vfull[slot], (uint32_t)(rows * HS * 4));
            bulk_g2s(kring + (size_t)slot * KR * KROW, kc + (size_t)t0 * HS, (uint32_t)(rows * HS * 4), &vfull[slot]);
        }
        __syncwarp();
    };
    if (bigk) {
        if (lane == 0) kprog[warp] = 0u;
        if (warp == 0) {
            fence_proxy_async();                            // the ring held weights the generic proxy has just read
            asm volatile("fence.proxy.async.global;" ::: "memory");      // the K rows were written through the generic proxy
            for (int c = 0; c < min(NKC, nkch); ++c) issue_k(c);
        }
    }
    constexpr int UU = 3;                      // key batches per round trip: 96 keys per CTA (6 batches in the long-context variant measured slower: 122 vs 138 tokens/s at the 13B shape)
    float4 kv[UU][EPL / 4];
    auto load_k = [&](int base) {
#pragma unroll
        for (int u = 0; u < UU; ++u) {
            const int t = base + (u * kConsumerWarps + warp) * 4 + rr;
            if (t < pos && t < te) {
                const float4* kp = reinterpret_cast<const float4*>(kc + (size_t)t * HS) + j;
#pragma unroll
                for (int q = 0; q < EPL / 4; ++q) kv[u][q] = __ldcg(kp + 8 * q);
            } else {
#pragma unroll
                for (int q = 0; q < EPL / 4; ++q) kv[u][q] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    };
    if constexpr (!bigk) load_k(tb);

    // ---- q, k, v rows of this head: 3 * HS tagged words, one 16-byte load (= one RoPE pair) per thread
    // RoPE + KV append (rope_v2 tf_operators.cpp:355-402; transformer.cpp:431-439)
    if (tid < 3 * (HS / 2)) {
        const int sect = tid / (HS / 2), i = tid - sect * (HS / 2);
        // sequence_rope_v2 (tensor.h:262-270) walks all bs*hgs rows of the q tensor with position pos0 + row: query head g
        // of a GQA group is rotated at pos + g*bs (== pos when n_heads == n_kv_heads).  Reproduced, not fixed.
        const float2 cs2 = __ldg(reinterpret_cast<const float2*>(p.rope) + (size_t)(sect == 0 ? pos + g * bs : pos) * (HS / 2) + i);
        const uint2* src = sv.qkvt + (sect == 0 ? (size_t)qh * HS : sect == 1 ? (size_t)dim + (size_t)kvh * HS : (size_t)dim + kv_dim + (size_t)kvh * HS) + 2 * i;
        uint4 w = ld_tag2(src);
        while (w.y != tag_qkv || w.w != tag_qkv) w = ld_tag2(src);
        const float x0 = __uint_as_float(w.x), x1 = __uint_as_float(w.z);
        if (sect == 0) {
            float o0, o1;
            rope_pair(cs2.x, cs2.y, x0, x1, o0, o1);
            q_s[2 * i] = o0; q_s[2 * i + 1] = o1;
        } else if (sect == 1) {
            float o0, o1;
            rope_pair(cs2.x, cs2.y, x0, x1, o0, o1);
            k_s[2 * i] = o0; k_s[2 * i + 1] = o1;
            if (g == 0 && part == 0) {
                float* krow = kc + (size_t)pos * HS;
                krow[k_cache_index(2 * i)] = o0;
                krow[k_cache_index(2 * i + 1)] = o1;
            }
        } else {
            v_s[2 * i] = x0; v_s[2 * i + 1] = x1;
            if (g == 0 && part == 0) {
                float* vblk = sv.vc + cache_off + (size_t)((2 * i) / DW) * p.max_seq * DW + ((size_t)(pos >> 2) * DW + (2 * i) % DW) * 4 + (pos & 3);
                vblk[0] = x0; vblk[4] = x1;
            }
        }
    }
    consumer_sync();
    pf.stop(tid, 10);
    pf.log(lane, warp, 16, 0);
    pf.mark(tid, 30, pf.trace_slot >= 0);

    // ---- scores for this part's keys: float dot_product_avx256 (x86_simd.cpp:1447-1468): 8 FMA chains, then 0 + l0 + ... + l7
    uint2* att_g = sv.score_t + (size_t)qh * p.score_stride;
    float mloc = -INFINITY;                        // running maximum of the scores this thread stores (softmax's max, fused)
    if constexpr (bigk) {
        const KSweepOut ko = k_sweep_ring<HS>(kring, vfull, kprog, kc, q_s, k_s, att, att_g, tb, kc_end, te, pos, NKC, nkch, cph, p.attn_scale, tag_score, ph, tid);
        ph = ko.ph; mloc = ko.mloc;
    } else {
        const float* qj = q_s + j;                 // q values of this AVX lane are re-read from shared memory (registers are scarce)
#pragma unroll 1
        for (int base = tb; base < te; base += kConsumerWarps * 4 * UU) {
#pragma unroll
            for (int u = 0; u < UU; ++u) {
                const int t = base + (u * kConsumerWarps + warp) * 4 + rr;
                float acc = 0.0f;
                if (t == pos) {
#pragma unroll
                    for (int q = 0; q < EPL / 4; ++q)
                        kv[u][q] = make_float4(k_s[8 * (4 * q) + j], k_s[8 * (4 * q + 1) + j], k_s[8 * (4 * q + 2) + j], k_s[8 * (4 * q + 3) + j]);
                }
#pragma unroll
                for (int q = 0; q < EPL / 4; ++q) {
                    acc = __fmaf_rn(kv[u][q].x, qj[8 * (4 * q)], acc);
                    acc = __fmaf_rn(kv[u][q].y, qj[8 * (4 * q + 1)], acc);
                    acc = __fmaf_rn(kv[u][q].z, qj[8 * (4 * q + 2)], acc);
                    acc = __fmaf_rn(kv[u][q].w, qj[8 * (4 * q + 3)], acc);
                }
                float tot = 0.0f;
#pragma unroll
                for (int k = 0; k < 8; ++k) tot = __fadd_rn(tot, __shfl_sync(kFull, acc, (rr << 3) + k));
                if (j == 0 && t < te) {
                    const float sc = __fmul_rn(tot, p.attn_scale);      // att.multiply(attn_scale), transformer.cpp:443
                    att[t] = sc;
                    mloc = fmaxf(mloc, sc);
                    if (cph > 1) st_tag(att_g + t, sc, tag_score);
                }
            }
            if (base + kConsumerWarps * 4 * UU < te) load_k(base + kConsumerWarps * 4 * UU);     // contexts beyond 32 * UU * cph keys: one more round trip per batch
        }
    }
    pf.stop(tid, 11);
    // ---- exchange: fetch the other parts' scores (two tagged words per load), waiting for their tag
    if (cph > 1) {
#pragma unroll 1
        for (int t = 2 * tid; t < n; t += 2 * kConsumerThreads) {
            const bool need0 = (t < tb || t >= te), need1 = (t + 1 < n) && (t + 1 < tb || t + 1 >= te);
            if (need0 || need1) {
                uint4 w = ld_tag2(att_g + t);
                while ((need0 && w.y != tag_score) || (need1 && w.w != tag_score)) w = ld_tag2(att_g + t);
                if (need0) { att[t] = __uint_as_float(w.x); mloc = fmaxf(mloc, __uint_as_float(w.x)); }
                if (need1) { att[t + 1] = __uint_as_float(w.z); mloc = fmaxf(mloc, __uint_as_float(w.z)); }
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mloc = fmaxf(mloc, __shfl_xor_sync(kFull, mloc, o));
    if (lane == 0) red[warp] = mloc;
    consumer_sync();
    if (gate && !big && tid == 0) st_shared_volatile_u32(gate, gate_val);       // q / k / v and the scores are in: Wo's prefetch may start (softmax and PV give it time)
    if (bigk && pvt == 0) {
        // the K sweep is over in every warp: the ring takes the first wave of V chunks (softmax gives them time to land)
        fence_proxy_async();
        for (int c = 0; c < min(NCH, n_chunks); ++c) issue_v(c);
    }
    pf.stop(tid, 12);
    pf.log(lane, warp, 13, 0);

    // ---- softmax_sisd (tf_operators.cpp:176-186): max (every score passed through exactly one thread's mloc above),
    //      expf(x - max), serial sum, divide
    float m = red[0];
#pragma unroll
    for (int w = 1; w < kConsumerWarps; ++w) m = fmaxf(m, red[w]);
#pragma unroll 1
    for (int t = tid; t < n; t += kConsumerThreads) att[t] = expf_exact(__fsub_rn(att[t], m));
    if (tid < 32) att[n + tid] = 0.0f;                     // the chains below read whole float4s / whole 32-row chunks of weights
    consumer_sync();
    if (RX) {
        // relaxed: per-thread partial sums, warp tree, 8 partials in order (not the reference's serial order)
        float part = 0.0f;
#pragma unroll 1
        for (int t = tid; t < n; t += kConsumerThreads) part = __fadd_rn(part, att[t]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part = __fadd_rn(part, __shfl_xor_sync(kFull, part, o));
        if (lane == 0) red[8 + warp] = part;
        consumer_sync();
        if (tid == 0) {
            float tot = red[8];
#pragma unroll
            for (int w = 1; w < kConsumerWarps; ++w) tot = __fadd_rn(tot, red[8 + w]);
            red[16] = tot;
        }
    } else if (tid == kSerialWarp * 32) {
        // one FP32 add chain in index order; loads run one batch ahead of the adds
        const float4* a4 = reinterpret_cast<const float4*>(att);
        const int nv = n >> 2;
        float sum = 0.0f;
        float4 c0 = a4[0], c1 = a4[1];
        int i = 0;
#pragma unroll 1
        for (; i + 2 <= nv; i += 2) {
            const float4 n0 = a4[i + 2], n1 = a4[i + 3];
            sum = __fadd_rn(sum, c0.x); sum = __fadd_rn(sum, c0.y); sum = __fadd_rn(sum, c0.z); sum = __fadd_rn(sum, c0.w);
            sum = __fadd_rn(sum, c1.x); sum = __fadd_rn(sum, c1.y); sum = __fadd_rn(sum, c1.z); sum = __fadd_rn(sum, c1.w);
            c0 = n0; c1 = n1;
        }
#pragma unroll 1
        for (int t = 4 * i; t < n; ++t) sum = __fadd_rn(sum, att[t]);
        red[16] = sum;
    }
    consumer_sync();
    const float sum = red[16];
#pragma unroll 1
    for (int t = tid; t < n; t += kConsumerThreads) {
        const float w = __fdiv_rn(att[t], sum);
        // the new token's weight goes aside (its V row is not in the staged cache): every weight from `pos` to the end of the
        // last chunk reads 0 and is skipped by the threshold test below
        if (t == pos) { red[17] = w; att[t] = 0.0f; } else att[t] = w;
    }
    consumer_sync();
    pf.stop(tid, 13);
    pf.log(lane, warp, 14, 0);

    // ---- weighted_sum (tf_operators.cpp:325-350): o = V[0]*w0; t >= 1: if |w_t| > 1e-15: o = fma(V[t], w_t, o) — one chain
    // per head dim, run by the PV group without CTA-wide synchronisation; its first thread refills the chunk ring.
    // The chain is one warp per 32 head dims and a dependent FMA issues every 4 cycles, so everything else must stay out
    // of its way: chunks are 32 rows = 8 quads of 4 positions ([t/4][DW][4] layout: one LDS.128 brings 4 rows of a dim, one
    // broadcast LDS.128 their 4 weights), the quads of chunk c+1 are loaded into the registers chunk c's quads just left,
    // and every row takes the same predicated FMA (rows from `pos` on carry weight 0; row 0, which the reference multiplies
    // unconditionally, is taken out of quad 0 before the loop).
    if (pvt >= 0) {
        float o = 0.0f;
        if (n_chunks > 0) {
            uint32_t slot = 0u;
            const int qstride = DW * 4;                     // floats between two quads of one head dim
            float4 va[8], wa[8];
            mbar_wait(&vfull[0], ph & 1u);
            ph ^= 1u;
            pf.stop(tid, 14);                               // waiting for the first V chunk
            {
                const float* vs = v_stage + (size_t)slot * VR * DW + pvt * 4;
#pragma unroll
                for (int u = 0; u < 8; ++u) { va[u] = *reinterpret_cast<const float4*>(vs + u * qstride); wa[u] = *reinterpret_cast<const float4*>(att + 4 * u); }
            }
            o = __fmul_rn(va[0].x, wa[0].x);
            wa[0].x = 0.0f;
#pragma unroll 1
            for (int c = 0; c < n_chunks; ++c) {
                const bool more = c + 1 < n_chunks;
                uint32_t nslot = slot + 1u;
                if (nslot == (uint32_t)NCH) nslot = 0u;
                const float* vs = v_stage + (size_t)nslot * VR * DW + pvt * 4;
                const float* ws = att + (c + 1) * VR;
                if (more) { mbar_wait(&vfull[nslot], (ph >> nslot) & 1u); ph ^= 1u << nslot; }
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    if (fabsf(wa[u].x) > 1e-15f) o = __fmaf_rn(va[u].x, wa[u].x, o);
                    if (fabsf(wa[u].y) > 1e-15f) o = __fmaf_rn(va[u].y, wa[u].y, o);
                    if (fabsf(wa[u].z) > 1e-15f) o = __fmaf_rn(va[u].z, wa[u].z, o);
                    if (fabsf(wa[u].w) > 1e-15f) o = __fmaf_rn(va[u].w, wa[u].w, o);
                    if (more) { va[u] = *reinterpret_cast<const float4*>(vs + u * qstride); wa[u] = *reinterpret_cast<const float4*>(ws + 4 * u); }
                }
                if (c + NCH < n_chunks) {
                    // chunk c has been consumed by every thread of the group: refill its slot with chunk c + NCH
                    if (DW > 32) asm volatile("bar.sync 2, %0;" :: "r"(DW) : "memory"); else __syncwarp(DW == 32 ? kFull : (kFull << (32 - DW)));      // the PV group is the top DW lanes of its warp
                    if (pvt == 0) issue_v(c + NCH);
                }
                slot = nslot;
            }
        }
        {   // the new token's row
            const float w = red[17], v = v_s[d0 + pvt];
            if (n_chunks == 0) o = __fmul_rn(v, w);
            else if (fabsf(w) > 1e-15f) o = __fmaf_rn(v, w, o);
        }
        delay_point<3>();
        st_tag(sv.attnt + (size_t)qh * HS + d0 + pvt, o, tag_out);
        if (pvt == 0) {
            *vphase = ph;
#ifdef FL_PROFILE
            if (pf.p && pf.trace_slot >= 0) pf.p[31] = gtimer();
#endif
        }
    }
    pf.stop(tid, 15);
    if (!(p.debug_skip & 4)) consumer_sync();     // the other warps start polling for the next phase only now: their strong loads share the LSU with the PV warp's shared-memory loads
    if (big && tid == 0) st_shared_volatile_u32(gate, gate_val);      // the ring is the producer's again
}

// ---------------------------------------------------------------------------------------------- the kernel
template <int QT, int GS, int HS, bool MS = false, bool RX = false, bool LC = false>
__global__ void __launch_bounds__(kMegaThreads, 1) decode_megakernel(const __grid_constant__ MegaParams p) {
    using RK = Rk<QT, GS>;
    const int n_seqs = MS ? p.n_seqs : 1;
    extern __shared__ __align__(16) uint8_t smem[];
    uint8_t* ring = smem + p.off_ring;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + p.off_bars);
    uint64_t* empty = full + p.n_slots;
    uint32_t* issued = reinterpret_cast<uint32_t*>(smem + p.off_misc) + 30;      // stages issued so far (producer -> consumers)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_slots = p.n_slots;

    if (tid == 0) {
        for (int i = 0; i < n_slots; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        uint64_t* vfull = reinterpret_cast<uint64_t*>(smem + p.off_vbars);
        for (int i = 0; i < 32; ++i) mbar_init(&vfull[i], 1);
        *issued = 0u;
        reinterpret_cast<uint32_t*>(smem + p.off_misc)[28] = 0u;      // V barrier phases
        reinterpret_cast<uint32_t*>(smem + p.off_misc)[24] = 0u;      // pair buffer 0 / 1: times released by the chain warp
        reinterpret_cast<uint32_t*>(smem + p.off_misc)[25] = 0u;
        reinterpret_cast<uint32_t*>(smem + p.off_misc)[31] = 0u;
        reinterpret_cast<uint32_t*>(smem + p.off_misc)[29] = 0u;      // phases the chain warp has finished
        reinterpret_cast<uint32_t*>(smem + p.off_misc)[20] = 0u;      // prefetch gate: phases (counted over all steps) the producer may stream
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const int n_phases = 4 * p.n_layers + 1;
    {
        const uint8_t** psrc = reinterpret_cast<const uint8_t**>(smem + p.off_psrc);
        for (int pi = tid; pi < n_phases; pi += kMegaThreads) psrc[pi] = phase_stream(p, pi);
        if (tid >= 64 && tid < 69) fill_geometry<QT, GS>(p, reinterpret_cast<int*>(smem + p.off_geom), tid - 64);
    }
    const int* geom = reinterpret_cast<const int*>(smem + p.off_geom);
    __syncthreads();
#ifndef FL_NO_SETMAXNREG
    if (warp >= kConsumerWarps) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" :: "n"(kRegsService));
        if (warp >= kConsumerWarps + 2) return;
    }
#endif

    if (warp == kConsumerWarps) {
        // ================= TMA producer =================
#ifdef FL_OLD_PRODUCER
        if (lane == 0) {
#else
        {       // the whole warp walks the schedule (warp-uniform control flow), one elected lane issues: see issue_stage_elect
#endif
            uint32_t sc = 0, slot = 0, par = 1;         // empty[] parity to wait for: 1 passes on a fresh barrier
            // In-flight window: stage i is issued only after stage i - window has LANDED.  The ring is deep so that it can
            // buffer microseconds of weights, but requests queued in the memory system are pure latency for everyone else
            // (Little's law: 148 SMs x 180 KB outstanding = 4 us at HBM speed) - and "everyone else" is the tagged-word
            // exchange every phase waits for.  ~40 KB in flight per SM already saturates HBM.
            uint32_t wslot = 0, wpar = 0;
            const uint32_t window = (uint32_t)p.window;
#pragma unroll 1
            for (int step = 0; step < p.n_steps; ++step) {
#pragma unroll 1
                for (int pi = 0; pi < n_phases; ++pi) {
                    const int* pg = geom + ((pi == n_phases - 1) ? 4 : (pi & 3)) * kGeomStride;
                    const int n_stages = pg[PG_NKC] * pg[PG_TT];      // per tile
                    // Prefetch gate.  Between two drains every CTA polls the tagged words of its next input, and those loads
                    // queue behind whatever the 148 producers have in flight (Little's law: 148 x 175 KB = 4 us at HBM speed).
                    // The ring refill for the NEXT phase is not needed before that input has arrived, so it is held back
                    // until the consumers say so (after their poll), and then runs during the rebuild's arithmetic.
                    if (!MS && !(p.debug_skip & (8 | 32))) {
                        const uint32_t need = (uint32_t)(step * n_phases + pi) + 1u;
                        const uint32_t* gate = reinterpret_cast<const uint32_t*>(smem + p.off_misc) + 20;
                        while ((int)(ld_shared_volatile_u32(gate) - need) < 0) __nanosleep(FL_GATE_SLEEP);
                    }
#pragma unroll 1
                    for (int sq = 0; sq < n_seqs; ++sq) {
                    if (MS && !(p.debug_skip & (8 | 32))) {
                        // several sequences: every (phase, sequence) pass is released by that sequence's own poll
                        const uint32_t need = (uint32_t)((step * n_phases + pi) * n_seqs + sq) + 1u;
                        const uint32_t* gate = reinterpret_cast<const uint32_t*>(smem + p.off_misc) + 20;
                        while ((int)(ld_shared_volatile_u32(gate) - need) < 0) __nanosleep(FL_GATE_SLEEP);
                    }
                    const uint8_t* src = reinterpret_cast<const uint8_t* const*>(smem + p.off_psrc)[pi];      // the stream is laid out in issue order
#pragma unroll 1
                    for (int t = 0; t < pg[PG_NT]; ++t) {
                        const uint32_t bytes = (uint32_t)RK::stage_bytes(pg[PG_LR + t + 1] - pg[PG_LR + t]);
#pragma unroll 1
                        for (int st = 0; st < n_stages; ++st) {
                            if (window < (uint32_t)n_slots && sc >= window) {
                                mbar_wait_sleep(&full[wslot], wpar);
                                if (++wslot == (uint32_t)n_slots) { wslot = 0; wpar ^= 1u; }
                            }
                            mbar_wait_sleep(&empty[slot], par);
#ifdef FL_OLD_PRODUCER
                            mbar_arrive_expect_tx(&full[slot], bytes);
                            bulk_g2s(ring + (size_t)slot * p.slot_bytes, src, bytes, &full[slot]);
                            __threadfence_block();                      // the barrier is armed before the count says so
                            st_shared_volatile_u32(issued, ++sc);
#else
                            issue_stage_elect(ring + (size_t)slot * p.slot_bytes, src, bytes, &full[slot], issued, ++sc);
#endif
                            src += bytes;
                            if (++slot == (uint32_t)n_slots) { slot = 0; par ^= 1u; }
                        }
                    }
                    }
                }
            }
        }
        return;
    }

    float2* pairbuf = reinterpret_cast<float2*>(smem + p.off_pairs);            // [2 buffers][kPairGroups][32 lanes]
    uint32_t* freed = reinterpret_cast<uint32_t*>(smem + p.off_misc) + 24;      // [2] releases of each pair buffer (chain warp -> consumers)

    if (warp == kConsumerWarps + 1) {
        // ================= chain warp =================
        // Follows the consumers' schedule superblock by superblock; lane i owns row i of the current tile.
        uint32_t sbseq = 0;                                  // superblocks so far (buffer = sbseq & 1)
        uint32_t phases_done = 0;
        Prof pf;
        pf.p = p.prof ? p.prof + (size_t)blockIdx.x * 32 : nullptr; pf.t0 = 0ull; pf.trace_slot = -1; pf.ev = nullptr;
        pf.evn = reinterpret_cast<unsigned int*>(smem + p.off_misc) + 31;
#pragma unroll 1
        for (int step = 0; step < p.n_steps; ++step) {
            const uint32_t tbase = p.epoch + 1u + (uint32_t)step * (uint32_t)(p.n_layers + 1) * kTagsPerLayer;
#pragma unroll 1
            for (int pi = 0; pi < n_phases; ++pi) {
                const int layer = pi >> 2, pk = (pi == n_phases - 1) ? 4 : (pi & 3);
                const int* pg = geom + pk * kGeomStride;
                struct { int tt, M; } ph = {pg[PG_TT], pg[PG_M]};
                pf.ev = ((step == p.n_steps - 1) && (layer == p.n_layers / 2) && pk < 4 && p.evlog && (p.debug_skip & 16) && (blockIdx.x == 7 || blockIdx.x == gridDim.x - 3)) ? p.evlog + (blockIdx.x == 7 ? 0 : 4096) : nullptr;
                const uint32_t tl = tbase + (uint32_t)layer * kTagsPerLayer;
#pragma unroll 1
                for (int sq = 0; sq < n_seqs; ++sq) {
                const SeqView sv = seq_view<MS>(p, sq);
                uint2* out = (pk == 0) ? sv.qkvt : (pk == 2) ? sv.hdt : sv.x1t;
                const uint32_t tag_out = tl + ((pk == 0) ? 1u : (pk == 1) ? 4u : (pk == 2) ? 5u : 6u);
                struct { int rb, nt; } pt = {pg[PG_RB], pg[PG_NT]};
                const int nkc = pg[PG_NKC], sk = pg[PG_SK], nsb = pg[PG_NSB];
                const int gstride = (kPairGroups / ph.tt) * 32;          // float2s per sub-stream in a pair buffer
                float best_v = -INFINITY;
                int best_i = 0x7fffffff;
#ifndef FL_OLD_REBUILD
                if (kRbMode != 2 && !RX && (pk == 0 || pk == 2 || pk == 4) && !(p.debug_skip & 8)) {
                    // the rmsnorm rebuild's sum of squares (build_activation): all 256 consumer threads have polled their words
                    // and stored them into the transposed vector
                    asm volatile("bar.sync 5, %0;" :: "n"(kConsumerThreads + 32) : "memory");
                    float* misc = reinterpret_cast<float*>(smem + p.off_misc);
                    if (lane == 0) {
                        // the input is here: the producer may prefetch again
                        const uint32_t gv = ld_shared_volatile_u32(reinterpret_cast<const uint32_t*>(misc) + 19);
                        if (gv) st_shared_volatile_u32(reinterpret_cast<uint32_t*>(misc) + 20, gv);
                    }
                    const float ss = sumsq_chain_t(reinterpret_cast<const float*>(smem + p.off_xt), p.dim, lane);
                    if (lane == 0) misc[18] = rms_scale(ss, p.dim);
                    pf.log(lane, 9, 23, 0);
                    asm volatile("bar.arrive 6, %0;" :: "n"(kConsumerThreads + 32) : "memory");
                }
#endif
#pragma unroll 1
                for (int t = 0; t < pt.nt; ++t) {
                    const int lr0 = pg[PG_LR + t], R = pg[PG_LR + t + 1] - lr0;
                    const int row = pt.rb + lr0 + lane;
                    const bool live = lane < R;
                    // residual input of this row (x1 += tmp, tensor.cpp:723): its word was validated by the consumers' earlier build
                    float x_old = 0.0f;
                    if ((pk == 1 || pk == 3) && live) x_old = __ldcg(reinterpret_cast<const float*>(sv.x1t + row));
                    float acc = 0.0f, acc2 = 0.0f;                     // acc2: the W3 row of the fused W1/W3 stream
#pragma unroll 1
                    for (int j = 0; j < nsb; ++j) {
                        int k0, S;
                        rk_superblock(nkc, sk, j, k0, S);
                        const int ng = S * RK::GPS;
                        const uint32_t buf = sbseq & 1u;
                        pf.log(lane, 9, 4, j);
                        asm volatile("bar.sync %0, %1;" :: "r"(3 + (int)buf), "n"(kConsumerThreads + 32) : "memory");     // the 8 consumers have arrived
                        pf.log(lane, 9, 5, j);
                        const float2* pb = pairbuf + (size_t)buf * kPairGroups * 32 + lane;
                        // the chain: loads run a batch ahead of the dependent FMAs
                        int g = 0;
#pragma unroll 1
                        for (; g + 8 <= ng; g += 8) {
                            float2 a[8], b[8];
#pragma unroll
                            for (int u = 0; u < 8; ++u) { a[u] = pb[(g + u) * 32]; if (ph.tt == 2) b[u] = pb[gstride + (g + u) * 32]; }
#pragma unroll
                            for (int u = 0; u < 8; ++u) { acc = __fmaf_rn(a[u].x, a[u].y, acc); if (ph.tt == 2) acc2 = __fmaf_rn(b[u].x, b[u].y, acc2); }
                        }
                        for (; g < ng; ++g) {
                            const float2 a = pb[g * 32];
                            acc = __fmaf_rn(a.x, a.y, acc);
                            if (ph.tt == 2) { const float2 b = pb[gstride + g * 32]; acc2 = __fmaf_rn(b.x, b.y, acc2); }
                        }
                        __syncwarp();
                        if (lane == 0) st_shared_volatile_u32(freed + buf, (sbseq >> 1) + 1u);        // buffer released
                        ++sbseq;
                    }
                    if (t == pt.nt - 1) delay_point<2>();
                    if (live) {
                        float v;
                        if (pk == 0 || pk == 4) v = acc;
                        else if (pk == 2) v = swiglu_exact(acc, acc2);
                        else v = __fadd_rn(x_old, acc);
                        if (pk == 4) {
                            if (!MS) p.logits[row] = v;        // fl_forward_batch returns token ids only
                            if (v > best_v || (v == best_v && row < best_i)) { best_v = v; best_i = row; }
                        } else {
                            st_tag(out + row, v, tag_out);
                        }
                    }
                    pf.log(lane, 9, 6, t);
                }
                __syncwarp();
                // Release side of the KV-cache hand-off: this CTA's consumer warps appended the token's K/V rows during attention
                // (plain stores, ordered before this point by the pair-buffer barriers).  The fence makes them visible
                // device-wide before the tagged rows this warp publishes next (hd, then x1), which every CTA polls before it
                // reads the cache for the next token.  Off the critical path: the Wo rows are already out.
#ifndef FL_NO_KV_FENCE
                if (pk == 1) __threadfence();
#endif
                if (lane == 0) st_shared_volatile_u32(reinterpret_cast<uint32_t*>(smem + p.off_misc) + 29, ++phases_done);      // pair buffers are idle until the next drain
                if (pk == 4) {
                    // per-CTA argmax partial (sampler.cpp:36-46: first index of the strict maximum)
                    const uint32_t tag_am = tbase + (uint32_t)p.n_layers * kTagsPerLayer + 1u;
                    float bv = best_v; int bi = best_i;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        const float ov = __shfl_xor_sync(kFull, bv, o);
                        const int oi = __shfl_xor_sync(kFull, bi, o);
                        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
                    }
                    if (lane == 0) st_relaxed_v4(sv.am + blockIdx.x, make_uint4(__float_as_uint(bv), tag_am, (uint32_t)bi, tag_am));
                }
                }
            }
        }
        return;
    }

    // ================= consumers =================
#ifndef FL_NO_SETMAXNREG
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" :: "n"(kRegsConsumer));
#endif
    uint8_t* xq = smem + p.off_xq;
    float* xs = reinterpret_cast<float*>(smem + p.off_xs);
    float* xt = reinterpret_cast<float*>(smem + p.off_xt);
    float* misc = reinterpret_cast<float*>(smem + p.off_misc);
    const uint4* xq4 = reinterpret_cast<const uint4*>(xq);
    const int n_attn_ctas = p.n_heads * p.cph;
    const bool attn_cta = (int)blockIdx.x < n_attn_ctas;
    const int my_head = blockIdx.x / p.cph, my_part = blockIdx.x % p.cph;
    uint32_t sc = 0, cseq = 0, sbseq = 0;      // stages / K chunks / superblocks so far
    uint32_t sb_sl = 0, sb_pr = 0;             // ring slot and parity of stage `sc`, advanced incrementally
    uint32_t phases_drained = 0;
    Prof pf;
    pf.p = p.prof ? p.prof + (size_t)blockIdx.x * 32 : nullptr;
    pf.t0 = pf.p ? (unsigned long long)clock64() : 0ull;
    pf.trace_slot = -1;
    pf.ev = nullptr;
    pf.evn = reinterpret_cast<unsigned int*>(smem + p.off_misc) + 31;

    // sequence state at launch: written by the previous kernel on this stream.  Kept in shared memory, not registers:
    // the phase loop below is register-bound (168 per thread with 9 warps on 4 schedulers) and must not spill.
    // several sequences: [4 * s + {0 token, 1 n_out, 2 pos at launch, 3 bs at launch}] in the free tail of the misc block
    int* sstate = reinterpret_cast<int*>(misc) + (MS ? 448 : 26);          // [0] next token, [1] n_out
    if (MS) {
        if (tid < n_seqs) { sstate[4 * tid] = p.st[tid].token; sstate[4 * tid + 1] = p.st[tid].n_out; sstate[4 * tid + 2] = p.st[tid].pos; sstate[4 * tid + 3] = p.st[tid].bs; }
    } else if (tid == 0) { sstate[0] = p.st->token; sstate[1] = p.st->n_out; }
    const int pos0 = p.st->pos, bs0 = p.st->bs;
    consumer_sync();

#pragma unroll 1
    for (int step = 0; step < p.n_steps; ++step) {
        const uint32_t tbase = p.epoch + 1u + (uint32_t)step * (uint32_t)(p.n_layers + 1) * kTagsPerLayer;   // tag(layer, k) = tbase + layer * 8 + k
        // embedding row (transformer.cpp:115-122): this CTA's slice of it becomes the input vector of layer 0
#pragma unroll 1
        for (int sq = 0; sq < n_seqs; ++sq) {
            const int token = sstate[MS ? 4 * sq : 0];
            const int e0 = (int)((long long)p.dim * blockIdx.x / gridDim.x), e1 = (int)((long long)p.dim * (blockIdx.x + 1) / gridDim.x);
            uint2* x1t = seq_view<MS>(p, sq).x1t;
            for (int i = e0 + tid; i < e1; i += kConsumerThreads) st_tag(x1t + i, __ldg(p.emb + (size_t)token * p.dim + i), tbase);
        }
        pf.stop(tid, 19);
#pragma unroll 1
        for (int pi = 0; pi < n_phases; ++pi) {
            const int layer = pi >> 2, pk = (pi == n_phases - 1) ? 4 : (pi & 3);
            const uint32_t tl = tbase + (uint32_t)layer * kTagsPerLayer;
            // tags: +0 layer input (= +6 of the previous layer), +1 qkv, +2 scores, +3 attention out, +4 x1 after Wo, +5 hd, +6 x1 after W2, +7 quantised hd
            const uint32_t tag_x_in = (layer == 0) ? tbase : tl - kTagsPerLayer + 6u;
#pragma unroll 1
            for (int sq = 0; sq < n_seqs; ++sq) {
            const SeqView sv = seq_view<MS>(p, sq);
            const int pos = (MS ? sstate[4 * sq + 2] : pos0) + step;
            // ---- the activation vector of this phase
            const uint2* in = (pk == 1) ? sv.attnt : (pk == 3) ? sv.hdt : sv.x1t;
            const uint32_t tag_in = (pk == 1) ? tl + 3u : (pk == 2) ? tl + 4u : (pk == 3) ? tl + 5u : tag_x_in;
            const float* gain = (pk == 0) ? p.att_norm + (size_t)layer * p.dim : (pk == 2) ? p.ffn_norm + (size_t)layer * p.dim : (pk == 4) ? p.out_norm : nullptr;
            const bool traced = (step == p.n_steps - 1) && (layer == p.n_layers / 2) && pk < 4;
            pf.trace_slot = traced ? 22 + pk : -1;
            pf.ev = (traced && p.evlog && (p.debug_skip & 16) && (blockIdx.x == 7 || blockIdx.x == gridDim.x - 3)) ? p.evlog + (blockIdx.x == 7 ? 0 : 4096) : nullptr;
            pf.log(lane, warp, 8, pk);          // build starts
            {
                const int K = (pk == 3) ? p.hidden : p.dim;
                // gate value after this poll (for the two-batch hd poll: after its first batch): this phase may be streamed; Wo is
                // released by the attention part once q / k / v are in (or at once where this CTA has no attention to do)
                uint32_t* gate = reinterpret_cast<uint32_t*>(smem + p.off_misc) + 20;
                const uint32_t gpi = MS ? (uint32_t)((step * n_phases + pi) * n_seqs + sq) : (uint32_t)(step * n_phases + pi);
                const bool no_attn_here = !(attn_cta && !(p.debug_skip & 8));
                const uint32_t gate_val = gpi + ((!MS && pk == 0 && no_attn_here) ? 2u : 1u);
                if (!(p.debug_skip & 8) && pk == 3) {
                    build_hd<QT, GS>(xq, xs, sv.hdt, sv.hdqt, tag_in, tl + 7u, K, tid, pf, gate, gate_val);
                } else if (!(p.debug_skip & 8)) {
                    // producers of the vector: every CTA owns dim * c / n rows of x1; the attention parts own HS / cph outputs each
                    build_activation<QT, GS, RX>(xq, xs, xt, misc, in, tag_in, gain, K, (pk == 1) ? n_attn_ctas : (int)gridDim.x,
                                                 (pk == 4 && blockIdx.x == 0) ? p.tap_norm : nullptr, tid, pf, (!MS && pk == 1) ? nullptr : gate, gate_val, phases_drained);
                }
            }
            pf.stop(tid, 1);
            pf.log(lane, warp, 7, pk);          // drain starts
            const int* pg = geom + pk * kGeomStride;
            struct { int tt, M; } ph = {pg[PG_TT], pg[PG_M]};
            // ---- drain this CTA's stages of the phase
            {
                struct { int nt; } pt = {pg[PG_NT]};
                const int nkc = pg[PG_NKC], sk = pg[PG_SK], nsb = pg[PG_NSB];
                const int gstride = (kPairGroups / ph.tt) * 32;          // float2s per sub-stream in a pair buffer
#ifdef FL_PROFILE
                if (pf.p && tid == kProfThread) atomicAdd(pf.p + 20, (unsigned long long)(ld_shared_volatile_u32(issued) - sc));   // stages the producer is ahead at drain start
#endif
                const uint32_t drain_sc0 = sc;
#pragma unroll 1
                for (int t = 0; t < pt.nt; ++t) {
                    const int R = pg[PG_LR + t + 1] - pg[PG_LR + t];
                    const bool live = lane < R;
#pragma unroll 1
                    for (int j = 0; j < nsb; ++j) {
                        int k0, S;
                        rk_superblock(nkc, sk, j, k0, S);
                        const uint32_t buf = sbseq & 1u;
                        // the chain warp has released this pair buffer (it is two superblocks behind at worst)
                        pf.stop(tid, 17);            // drain bookkeeping
                        while ((int)(ld_shared_volatile_u32(freed + buf) - (sbseq >> 1)) < 0) __nanosleep(20);
                        pf.stop(tid, 16);            // waiting for the chain warp to release the pair buffer
                        pf.log(lane, warp, 10, j);
                        float2* pb = pairbuf + (size_t)buf * kPairGroups * 32;
                        // my K chunks of this superblock: chunks are dealt round-robin over the warps, continuing across superblocks
                        uint32_t q = ((uint32_t)warp - cseq) & 7u;
                        uint32_t rel = q * (uint32_t)ph.tt;                              // position in issue order, relative to sc
                        uint32_t sl = sb_sl + rel, pr = sb_pr;
                        while (sl >= (uint32_t)n_slots) { sl -= n_slots; pr ^= 1u; }
#pragma unroll 1
                        for (; q < (uint32_t)S; q += 8) {
#pragma unroll 1
                            for (int m = 0; m < ph.tt; ++m) {
                                pf.log(lane, warp, 1, (int)rel + m);
                                while ((int)(ld_shared_volatile_u32(issued) - (sc + rel + (uint32_t)m)) <= 0) __nanosleep(FL_ISSUED_SLEEP);      // fact 4 in the header
                                mbar_wait(&full[sl], pr);
                                pf.stop(tid, (sc + rel + (uint32_t)m - drain_sc0 < 16u) ? 21 : 7);       // waiting for weights = the stream is the limit (21: the stages prefetched during the stall)
                                pf.log(lane, warp, 2, (int)rel + m);
                                if (live && !(p.debug_skip & 1)) stage_pairs<QT, GS>(ring + (size_t)sl * p.slot_bytes, R, xq4 + (k0 + (int)q) * RK::PIECES, xs + (k0 + (int)q) * RK::GPS, lane,
                                                             pb + m * gstride + (int)q * RK::GPS * 32);
                                __syncwarp();
                                if (lane == 0) mbar_arrive(&empty[sl]);
                                pf.stop(tid, 2 + pk);
                                pf.log(lane, warp, 3, (int)rel + m);
                                if (++sl == (uint32_t)n_slots) { sl = 0; pr ^= 1u; }
                            }
                            rel += 8u * (uint32_t)ph.tt;
                            sl += (8u - 1u) * (uint32_t)ph.tt;
                            while (sl >= (uint32_t)n_slots) { sl -= n_slots; pr ^= 1u; }
                        }
                        sc += (uint32_t)(S * ph.tt);
                        sb_sl += (uint32_t)(S * ph.tt);
                        while (sb_sl >= (uint32_t)n_slots) { sb_sl -= n_slots; sb_pr ^= 1u; }
                        cseq += (uint32_t)S;
                        ++sbseq;
                        asm volatile("bar.arrive %0, %1;" :: "r"(3 + (int)buf), "n"(kConsumerThreads + 32) : "memory");     // my pairs of this superblock are in the buffer
                    }
                }
            }
            ++phases_drained;
            pf.stop(tid, 2 + pk);
#ifdef FL_PROFILE
            if (pf.p && traced && lane == 0) atomicMax(pf.p + 26 + pk, gtimer());
#endif
            if (pk == 4) {
                // ---- argmax (sampler.cpp:36-46): the chain warps published one partial per CTA as tagged words; every CTA
                //      reduces all of them, so every CTA knows the next token without another round trip
                const uint32_t tag_am = tbase + (uint32_t)p.n_layers * kTagsPerLayer + 1u;
                if (warp == 0) {
                    float bv = -INFINITY; int bi = 0x7fffffff;
                    for (int i = lane; i < (int)gridDim.x; i += 32) {
                        uint4 w = ld_relaxed_v4(sv.am + i);
                        while (w.y != tag_am || w.w != tag_am) { __nanosleep(100); w = ld_relaxed_v4(sv.am + i); }
                        const float v = __uint_as_float(w.x); const int ix = (int)w.z;
                        if (v > bv || (v == bv && ix < bi)) { bv = v; bi = ix; }
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        const float ov = __shfl_xor_sync(kFull, bv, o);
                        const int oi = __shfl_xor_sync(kFull, bi, o);
                        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
                    }
                    if (lane == 0) {
                        if (bi == 0x7fffffff) bi = 0;
                        sstate[MS ? 4 * sq : 0] = bi;
                        if (blockIdx.x == 0) {
                            // sequence state for the host and the next launch
                            SeqState* st = p.st + (MS ? sq : 0);
                            p.argmax_out[MS ? sq : 0] = bi;
                            const int n_out = sstate[MS ? 4 * sq + 1 : 1];
                            if (n_out < p.out_cap) p.out_tokens[(MS ? (size_t)sq * p.out_cap : 0) + n_out] = bi;
                            st->n_out = n_out + 1; st->token = bi; st->pos = pos + 1; st->bs = 1;
                            sstate[MS ? 4 * sq + 1 : 1] = n_out + 1;
                        }
                    }
                }
                consumer_sync();
            }
            }       // sequences
            if (pk == 0 && attn_cta && !(p.debug_skip & 8)) {
                // ---- attention (transformer.cpp:136, :397-455), one sequence after the other
                consumer_sync();            // the V stage aliases the activation image the other warps may still be draining with
#pragma unroll 1
                for (int sq = 0; sq < n_seqs; ++sq) {
                    const int pos = (MS ? sstate[4 * sq + 2] : pos0) + step;
                    const int bs = step == 0 ? (MS ? sstate[4 * sq + 3] : bs0) : 1;
                    attention_part<HS, RX, LC>(p, seq_view<MS>(p, sq), smem, layer, my_head, my_part, pos, bs, tl + 1u, tl + 2u, tl + 3u, phases_drained, tid, pf,
                                       MS ? nullptr : reinterpret_cast<uint32_t*>(smem + p.off_misc) + 20, (uint32_t)(step * n_phases + pi) + 2u);
                }
            }
        }
        pf.stop(tid, 18);
    }
}

}  // namespace fl
