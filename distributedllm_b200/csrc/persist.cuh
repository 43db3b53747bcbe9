// persist.cuh -- the single-token step of a whole slice as ONE persistent kernel (exact mode).
//
// The multi-kernel step (kernels.cuh: 4 launches per layer + attention) spends ~55 % of its time in the five dependency
// edges of a layer: at every kernel boundary the weight stream stops (the next kernel's CTAs cannot become resident
// before the previous kernel's CTAs free their shared memory), the consumer prologue pays an L2 round trip under a
// saturated memory system, and the narrow matrices (wo, w2) start cold.  Here ONE kernel runs all layers:
//
//   * grid = one CTA per SM, 16 consumer warps (4 GROUPS of 4 warps) + 1 producer warp;
//   * the CTA owns ONE ring of shared-memory stages (all the shared memory that is left, ~184 KB) fed by one producer
//     lane in a fixed global order: layer by layer, matrix by matrix (qkv, wo, w1|w3, w2), and inside a matrix the
//     stages of the CTA's active groups interleaved round-robin.  Whatever is being consumed therefore has the WHOLE
//     ring in flight -- one group on a narrow matrix (wo, w2) as much as three groups on a wide one -- and the ring
//     runs ahead into the next matrix / layer while the consumers sit in a dependency.  The producer never waits for
//     an activation (weights do not depend on them): the HBM stream runs THROUGH every dependency of the layer
//     (RMSNorm, attention, SiLU gate) and is only ever throttled by ring space;
//   * consumers walk the same schedule.  Phases are separated by grid-wide progress counters in global memory
//     (release: fence + atomicAdd by the group that finished its tiles; acquire: one polling thread per CTA + CTA
//     barrier), not by kernel boundaries;
//   * attention runs inside the kernel, one head per CTA (CTAs 0..H-1; the other CTAs keep prefetching), with the
//     arithmetic of k_attn128 -- the four cluster ranks become the four warp-quads of one CTA and the distributed
//     shared-memory exchange becomes plain shared memory;
//   * the per-block arithmetic (dp4a -> fadd -> fma in block order, 8 lanes, fixed hsum), the prologues (RMSNorm * w ->
//     Q8_0) and the epilogues (+residual, SiLU gate -> Q8_0) are those of k_gemv, bit for bit.
// Reference arithmetic: ggml.c:2431-2455 (dot), 1215-1252 (Q8_0), 10309-10352 (RMSNorm), 11956-12055 (RoPE),
// 11524-11590 (softmax), 2323-2357 (f16 dot); graph order tensor_processor.cpp:537-766.
#pragma once
#include "kernels.cuh"

namespace b200 {

constexpr int kPGroups = 4;                            // consumer groups per CTA
constexpr int kPConsumers = kPGroups * kConsumers;     // 512 consumer threads
constexpr int kPThreads = kPConsumers + 32;            // + producer warp (warp 16)
constexpr int kPPhases = 5;                            // counters per layer: qkv, attention, wo, w1|w3, w2 (= next layer's input)
constexpr int kPTraceSlots = 16;

__host__ __device__ constexpr int p_slot_bytes(int wt) { return 16 * chunk_bytes(wt); }   // 4 quads x 4 row-groups

struct PMat {                 // one weight matrix as the persistent kernel walks it
    const uint8_t * data;
    int n_tiles, nbq, TR, sq; // row-groups per tile; quads per ring stage (sq * TR * chunk <= slot)
    long long tile_bytes;
};

struct PLayer {
    PMat qkv, wo, w13, w2;
    const float * attn_norm, * ffn_norm;
    const float * x_in; float * x_out;     // layer input / output row [E]
    uint16_t * kc, * vc;                   // this layer's cache rows [n_ctx][E] of the current session
};

struct PersistArgs {
    const PLayer * layers; int L;
    int E, FF, H, n_ctx, nb_E, nbqE, nbqF;
    const int * n_past;
    float * qkv, * att, * ffin;
    int * aq_att; float * da_att; int * aq_gate; float * da_gate; float dscale;
    const float2 * cs; const uint16_t * texp, * tsilu;
    int * cnt;                             // [L][kPPhases], zero before the launch
    float kq_scale;
    int NS;                                // ring slots of the CTA
    unsigned long long * trace;            // optional [cta][L][kPTraceSlots]
};

// shared-memory carve-up, identical on host and device
struct PSmem {
    size_t ring, a_s, da_s, scratch, gq, bars, red, total;
};
__host__ __device__ inline PSmem p_smem_layout(int wt, int NS, int nbq_max, int E, int n_ctx) {
    PSmem m;
    size_t off = 0;
    m.ring = off;    off += (size_t) NS * p_slot_bytes(wt);
    m.a_s = off;     off += (size_t) nbq_max * 128;
    m.da_s = off;    off += (size_t) nbq_max * 16;
    // attention scratch: scores f32 [n_ctx] | p16 [n_ctx] | partials [4][8][128] f32 | q,k,v rows fp16
    size_t att = (((size_t)((n_ctx + 3) & ~3) * 4 + (size_t)((n_ctx + 7) & ~7) * 2) + 15) & ~(size_t) 15;
    att += 4 * 8 * 128 * 4 + 3 * 256;
    m.scratch = off; off += att;
    m.gq = off;      off += (size_t) kPGroups * 32 * 4;
    m.red = off;     off += 16 * 8 + 16 * 4 + 64;
    off = (off + 7) & ~(size_t) 7;
    m.bars = off;    off += (size_t) 2 * NS * 8;
    m.total = off;
    return m;
}

__device__ __forceinline__ int ld_acquire_gpu(const int * p) {
    int v; asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ int ld_relaxed_gpu(const int * p) {
    int v; asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
// release-add: the group's result stores (made visible to the executing thread by the group barrier before it) are
// ordered before the counter update at gpu scope -- one instruction instead of membar.gl in every thread + atomicAdd
__device__ __forceinline__ void red_release_gpu(int * p, int v) {
    asm volatile("red.release.gpu.global.add.s32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

#define P_TRACE(slot) do { if (a.trace && tid == 0) a.trace[((size_t) blockIdx.x * a.L + il) * kPTraceSlots + (slot)] = gtime(); } while (0)

// ---- ring schedule: `idx0` = global stage index at which the current matrix starts; advanced identically by the
// producer lane and by every consumer thread of the CTA.  Stage `st` of the j-th active group of a round sits at global
// index idx0 + st * n_act + j, in slot (index % NS), at the slot's (index / NS)-th use.
struct RingCur { int idx0; long long wait_clk; };

// number of groups of this CTA that own a tile of W in round `round` (groups are filled in order: 0, 1, ...)
__device__ __forceinline__ int p_active_groups(int n_tiles, int cta, int n_cta, int round) {
    const int left = n_tiles - round * kPGroups * n_cta - cta;          // tiles of this round at or after this CTA's column
    if (left <= 0) return 0;
    const int n = (left + n_cta - 1) / n_cta;
    return n < kPGroups ? n : kPGroups;
}

// One weight matrix: the tiles of this group, NC = 1.  EPI_RESID: y[row] = dot + resid[row];  EPI_STORE: y[row] = dot;
// EPI_GATEQ: g = silu(w1 x) * (w3 x), 32 gate rows of a tile = one Q8_0 block of w2's input, quantised by the group.
template <int WT, int G, int EPI>            // TR = 4 G row-groups per tile, 4 / G quads per ring stage: one stage = one slot
__device__ __forceinline__ int p_run_matrix(const PersistArgs & a, const PMat & W, RingCur & rc, uint8_t * ring, uint64_t * full, uint64_t * empty,
                                            const int * a_s, const float * da_s, int cta, int n_cta, int wig, int lane,
                                            const float * resid, float * y, int out_rows, float * gq_g, int grp) {
    constexpr int CB = (WT == kWT_Q4_0) ? kQ4Chunk : kQ8Chunk;
    constexpr int SLOT = 16 * CB;
    constexpr int TR = kWPC * G, sq = kQS / G;
    const int NS = a.NS;
    const int n_stage = W.nbq / sq;
    constexpr bool active = true;
    const int r = lane >> 2, w = lane & 3;
    int done = 0;
    for (int round = 0; ; round++) {
        const int n_act = p_active_groups(W.n_tiles, cta, n_cta, round);
        if (n_act == 0) break;
        const int tile = (round * kPGroups + grp) * n_cta + cta;
        if (grp >= n_act) { rc.idx0 += n_stage * n_act; continue; }
        float acc[G][2];
        #pragma unroll
        for (int g = 0; g < G; g++) { acc[g][0] = 0.f; acc[g][1] = 0.f; }
        int slot = (rc.idx0 + grp) % NS, use = (rc.idx0 + grp) / NS;
        for (int st = 0; st < n_stage; st++) {
            if (a.trace) { const long long c0 = clock64(); mbar_wait(&full[slot], use & 1); rc.wait_clk += clock64() - c0; }
            else mbar_wait(&full[slot], use & 1);
            if (active) {
                const uint8_t * base = ring + (size_t) slot * SLOT + (size_t)(wig * G) * CB;
                #pragma unroll
                for (int qi = 0; qi < sq; qi++) {
                    const int Q = st * sq + qi;
                    uint4 wv[G], wv2[G]; uint2 sc[G];
                    #pragma unroll
                    for (int g = 0; g < G; g++) {
                        const uint8_t * ch = base + (size_t)(qi * TR + g) * CB;
                        wv[g] = *(const uint4 *)(ch + lane * 16);
                        if (WT == kWT_Q8_0) { wv2[g] = *(const uint4 *)(ch + 512 + lane * 16); sc[g] = *(const uint2 *)(ch + 1024 + r * 8); }
                        else sc[g] = *(const uint2 *)(ch + 512 + r * 8);
                    }
                    const int4 * ap = (const int4 *)(a_s + Q * 32 + w * 8);
                    const int4 a01 = ap[0], a23 = ap[1];           // {lo0,hi0,lo1,hi1}, {lo2,hi2,lo3,hi3}
                    const float4 dav = *(const float4 *)(da_s + Q * 4);
                    const int alo[4] = {a01.x, a01.z, a23.x, a23.z};
                    const int ahi[4] = {a01.y, a01.w, a23.y, a23.w};
                    const float da[4] = {dav.x, dav.y, dav.z, dav.w};
                    #pragma unroll
                    for (int g = 0; g < G; g++) {
                        const uint32_t ww[4] = {wv[g].x, wv[g].y, wv[g].z, wv[g].w};
                        const uint32_t ww2[4] = {wv2[g].x, wv2[g].y, wv2[g].z, wv2[g].w};
                        const uint32_t sw[2] = {sc[g].x, sc[g].y};
                        #pragma unroll
                        for (int bq = 0; bq < 4; bq++) {
                            const uint16_t dh = (uint16_t)(sw[bq >> 1] >> (16 * (bq & 1)));
                            const float D = fmul(h2f(dh), da[bq]);
                            int lo, hi;
                            if (WT == kWT_Q4_0) { lo = (int)((ww[bq] << 4) & 0xF0F0F0F0u); hi = (int)(ww[bq] & 0xF0F0F0F0u); }
                            else                { lo = (int) ww[bq]; hi = (int) ww2[bq]; }
                            const float f0 = fadd(__int_as_float(__dp4a(lo, alo[bq], kMagicI)), -kMagic);
                            const float f1 = fadd(__int_as_float(__dp4a(hi, ahi[bq], kMagicI)), -kMagic);
                            acc[g][0] = ffma(D, f0, acc[g][0]);
                            acc[g][1] = ffma(D, f1, acc[g][1]);
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[slot]);
            slot += n_act; if (slot >= NS) { slot -= NS; use++; }
        }
        rc.idx0 += n_stage * n_act;
        // hsum_float_8 order ((a0+a4)+(a2+a6)) + ((a1+a5)+(a3+a7))
        float res[G];
        #pragma unroll
        for (int g = 0; g < G; g++) {
            float t = fadd(acc[g][0], acc[g][1]);
            t = fadd(t, __shfl_xor_sync(0xffffffffu, t, 2));
            t = fadd(t, __shfl_xor_sync(0xffffffffu, t, 1));
            res[g] = t;
        }
        if (EPI == EPI_GATEQ) {
            // tile = 8 row-groups: warp wig owns gate rows [8 wig, 8 wig + 8) of the tile's 32 (w1 group g = 0, w3 group g = 1)
            const int row = (tile * kWPC + wig) * 8 + r;
            if (w == 0) gq_g[wig * 8 + r] = row < out_rows ? fmul(h2f(a.tsilu[f2h(res[0])]), res[G - 1]) : 0.f;
            named_bar_sync(2 + grp, kConsumers);
            if (wig == 0) warp_quant_block(gq_g[lane], lane, a.aq_gate, a.da_gate, tile, a.dscale);
            named_bar_sync(2 + grp, kConsumers);
        } else if (active && w == 0) {
            #pragma unroll
            for (int g = 0; g < G; g++) {
                const int row = (tile * TR + wig * G + g) * 8 + r;
                if (row < out_rows) {
                    float v = res[g];
                    if (EPI == EPI_RESID) v = fadd(v, __ldcg(resid + row));
                    y[row] = v;
                }
            }
        }
        done++;
    }
    return done;
}

// RMSNorm * weight -> Q8_0 act-quant of the row x[K] into a_s / da_s: the arithmetic of k_gemv's fused prologue, spread
// over all 512 consumer threads -- FOUR threads per 32-value block (8 values each: coalesced 32-byte loads, the block's
// amax by two shuffles, two packed words per thread), up to two blocks per thread-quad (K <= 8192).
template <int WT>
__device__ __forceinline__ void p_pro_norm_quant(const float * x, const float * nw, int K, int nb, int nbq, int * a_s, float * da_s,
                                                 double * red, int tid) {
    const int warp = tid >> 5, lane = tid & 31, j = tid & 3;
    float v[2][8], wn[2][8];
    #pragma unroll
    for (int it = 0; it < 2; it++) {
        const int b = (tid >> 2) + 128 * it;
        const bool own = b < nb;
        #pragma unroll
        for (int h = 0; h < 2; h++) {
            const float4 t = own ? __ldcg((const float4 *)(x + b * 32 + j * 8 + h * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 u = own ? ldg_keep(nw + b * 32 + j * 8 + h * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
            v[it][h*4] = t.x; v[it][h*4+1] = t.y; v[it][h*4+2] = t.z; v[it][h*4+3] = t.w;
            wn[it][h*4] = u.x; wn[it][h*4+1] = u.y; wn[it][h*4+2] = u.z; wn[it][h*4+3] = u.w;
        }
    }
    for (int b = nb + tid; b < nbq * 4; b += kPConsumers) {            // padding blocks
        int * dst = a_s + (b >> 2) * 32 + (b & 3) * 2;
        for (int w = 0; w < 4; w++) { dst[w * 8] = 0; dst[w * 8 + 1] = 0; }
        da_s[b] = 0.f;
    }
    double s = 0.0;
    #pragma unroll
    for (int it = 0; it < 2; it++)
        #pragma unroll
        for (int e = 0; e < 8; e++) s += widen_nonneg(fmul(v[it][e], v[it][e]));
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) red[warp] = s;
    named_bar_sync(1, kPConsumers);
    double tot = 0.0;
    #pragma unroll
    for (int i = 0; i < 16; i++) tot += red[i];
    const float scale = __fdiv_rn(1.0f, __fsqrt_rn(fadd((float)(tot / (double) K), 1e-6f)));
    #pragma unroll
    for (int it = 0; it < 2; it++) {
        const int b = (tid >> 2) + 128 * it;
        const bool own = b < nb;
        float q[8];
        float amax = 0.f;
        #pragma unroll
        for (int e = 0; e < 8; e++) { q[e] = fmul(fmul(v[it][e], scale), wn[it][e]); amax = fmaxf(amax, fabsf(q[e])); }
        amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 1));
        amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, 2));
        if (own) {
            const float d = h2f(f2h(__fdiv_rn(amax, 127.f)));
            const float id = amax != 0.f ? __fdiv_rn(127.f, amax) : 0.f;
            int * dst = a_s + (b >> 2) * 32 + (b & 3) * 2;
            #pragma unroll
            for (int k = 0; k < 2; k++) {
                uint32_t pk = 0;
                #pragma unroll
                for (int e = 0; e < 4; e++) pk |= ((uint32_t)(rint_small(fmul(q[k*4 + e], id)) & 0xFF)) << (8 * e);
                const int ww = 2 * j + k;                          // 32-bit word of the block: values 4 ww .. 4 ww + 3
                dst[(ww & 3) * 8 + (ww >> 2)] = (int) pk;
            }
            if (j == 0) da_s[b] = (WT == kWT_Q4_0) ? fmul(d, 0.0625f) : d;
        }
    }
    named_bar_sync(1, kPConsumers);
}

// activation already quantised by its producer (attention / gate epilogue): copy words + scales into shared memory
__device__ __forceinline__ void p_pro_preq(const int * aq, const float * da, int nbq, int * a_s, float * da_s, int tid) {
    const int4 * s4 = (const int4 *) aq; int4 * d4 = (int4 *) a_s;
    for (int i = tid; i < nbq * 8; i += kPConsumers) d4[i] = __ldcg(s4 + i);
    const float4 * f4 = (const float4 *) da; float4 * g4 = (float4 *) da_s;
    for (int i = tid; i < nbq; i += kPConsumers) g4[i] = __ldcg(f4 + i);
    named_bar_sync(1, kPConsumers);
}

// grid-wide dependency: one thread polls the progress counter, the CTA barrier releases everybody
__device__ __forceinline__ void p_wait_counter(const int * cnt, int target, int tid) {
    if (tid == 0) { while (ld_relaxed_gpu(cnt) < target) { } (void) ld_acquire_gpu(cnt); }
    named_bar_sync(1, kPConsumers);
}

// Attention of head h for the token at position `pos` (T = pos + 1), the arithmetic of k_attn128<FUSE> with the four
// cluster ranks folded into one CTA of 512 threads.  Fused: RoPE of q / k, fp16 rounding, KV append of the new row,
// and the Q8_0 quantisation of the output for the wo matmul.
__device__ __forceinline__ void p_attention_head(const PersistArgs & a, const PLayer & Lw, int h, int pos, uint8_t * scratch, float * redf, double * redd, int tid) {
    const int E = a.E, tcount = pos + 1, T = pos + 1;
    const int warp = tid >> 5, lane = tid & 31;
    float * sc = (float *) scratch;
    uint16_t * p16 = (uint16_t *)(sc + ((a.n_ctx + 3) & ~3));
    float * partl = (float *)(scratch + ((((size_t)((a.n_ctx + 3) & ~3) * 4 + (size_t)((a.n_ctx + 7) & ~7) * 2) + 15) & ~(size_t) 15));
    uint16_t * q16s = (uint16_t *)(partl + 4 * 8 * 128), * k16s = q16s + 128, * v16s = k16s + 128;
    uint16_t * kc = Lw.kc, * vc = Lw.vc;
    // ---- phase 0: q, and the new k / v row
    if (tid < 64) {
        const float2 cs = a.cs[(size_t) pos * 64 + tid];
        const float * row = a.qkv + h * 128;
        const float2 q = __ldcg((const float2 *)(row + 2 * tid));
        const float2 k = __ldcg((const float2 *)(row + E + 2 * tid));
        const float2 v = __ldcg((const float2 *)(row + 2 * E + 2 * tid));
        const float q0 = fsub(fmul(q.x, cs.x), fmul(q.y, cs.y)), q1 = fadd(fmul(q.x, cs.y), fmul(q.y, cs.x));
        const float k0 = fsub(fmul(k.x, cs.x), fmul(k.y, cs.y)), k1 = fadd(fmul(k.x, cs.y), fmul(k.y, cs.x));
        const uint32_t qq = (uint32_t) f2h(q0) | ((uint32_t) f2h(q1) << 16);
        const uint32_t kk = (uint32_t) f2h(k0) | ((uint32_t) f2h(k1) << 16);
        const uint32_t vv = (uint32_t) f2h(v.x) | ((uint32_t) f2h(v.y) << 16);
        ((uint32_t *) q16s)[tid] = qq; ((uint32_t *) k16s)[tid] = kk; ((uint32_t *) v16s)[tid] = vv;
        *(uint32_t *)(kc + (size_t) pos * E + h * 128 + 2 * tid) = kk;
        *(uint32_t *)(vc + (size_t) pos * E + h * 128 + 2 * tid) = vv;
    }
    named_bar_sync(1, kPConsumers);
    // ---- phase 1: scores; 4 lanes per position, lane ql owns 16-byte vectors m = ql + 4c of the 256-byte key row
    {
        const int ql = tid & 3;
        float qf[4][8];
        #pragma unroll
        for (int c = 0; c < 4; c++)
            #pragma unroll
            for (int e = 0; e < 8; e++) qf[c][e] = h2f(q16s[32 * c + 8 * ql + e]);
        // two positions per thread per iteration: 8 x 16-byte loads in flight hide the (L2 / HBM) latency of the key rows
        for (int t0 = 0; t0 < tcount; t0 += kPConsumers / 2) {
            uint4 kv[2][4];
            bool valid[2];
            #pragma unroll
            for (int u = 0; u < 2; u++) {
                const int t = t0 + u * (kPConsumers / 4) + (tid >> 2);
                valid[u] = t < tcount;
                if (valid[u] && t != pos) {
                    const uint16_t * krow = kc + (size_t) t * E + h * 128;
                    #pragma unroll
                    for (int c = 0; c < 4; c++) kv[u][c] = __ldcg((const uint4 *)(krow + 32 * c + 8 * ql));
                } else {
                    #pragma unroll
                    for (int c = 0; c < 4; c++) kv[u][c] = *(const uint4 *)(k16s + 32 * c + 8 * ql);
                }
            }
            #pragma unroll
            for (int u = 0; u < 2; u++) {
                const int t = t0 + u * (kPConsumers / 4) + (tid >> 2);
                float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                #pragma unroll
                for (int c = 0; c < 4; c++) {
                    const uint32_t w4[4] = {kv[u][c].x, kv[u][c].y, kv[u][c].z, kv[u][c].w};
                    #pragma unroll
                    for (int e = 0; e < 8; e++) {
                        const uint16_t kh = (uint16_t)(w4[e >> 1] >> (16 * (e & 1)));
                        acc[e] = ffma(h2f(kh), qf[c][e], acc[e]);
                    }
                }
                float v8[8];
                #pragma unroll
                for (int e = 0; e < 8; e++) {                        // (x0 + x2) + (x1 + x3)
                    float x = fadd(acc[e], __shfl_xor_sync(0xffffffffu, acc[e], 2));
                    v8[e] = fadd(x, __shfl_xor_sync(0xffffffffu, x, 1));
                }
                const float u0 = fadd(v8[0], v8[4]), u1 = fadd(v8[1], v8[5]), u2 = fadd(v8[2], v8[6]), u3 = fadd(v8[3], v8[7]);
                const float dot = fadd(fadd(u0, u1), fadd(u2, u3));
                if (valid[u] && ql == 0) sc[t] = fmul(dot, a.kq_scale);
            }
        }
    }
    named_bar_sync(1, kPConsumers);
    // ---- phase 2: softmax
    float mx = -INFINITY;
    for (int t = tid; t < tcount; t += kPConsumers) mx = fmaxf(mx, sc[t]);
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) redf[warp] = mx;
    named_bar_sync(1, kPConsumers);
    mx = redf[0];
    #pragma unroll
    for (int i = 1; i < 16; i++) mx = fmaxf(mx, redf[i]);
    double s = 0.0;
    for (int t = tid; t < tcount; t += kPConsumers) {
        const float e = h2f(__ldg(a.texp + f2h(fsub(sc[t], mx))));
        sc[t] = e; s += (double) e;
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) redd[warp] = s;
    named_bar_sync(1, kPConsumers);
    double S = 0.0;
    #pragma unroll
    for (int i = 0; i < 16; i++) S += redd[i];                      // fp16-valued terms: exact in any order
    const float inv = (float)(1.0 / S);
    for (int t = tid; t < tcount; t += kPConsumers) p16[t] = f2h(fmul(sc[t], inv));
    named_bar_sync(1, kPConsumers);
    // ---- phase 3: V.p slot partials; quad g owns AVX vector j = g (positions 32k + 8g + l), thread (l, cg) 8 channels
    const int npT = T & ~31, lim = min(npT, tcount);
    {
        const int g = tid >> 7, l = (tid >> 4) & 7, cg = tid & 15;
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        // four value rows in flight per thread; the FMAs stay in position order (slot accumulators of ggml_vec_dot_f16)
        for (int tb = 8 * g + l; tb < lim; tb += 128) {
            uint4 vv[4];
            #pragma unroll
            for (int u = 0; u < 4; u++) {
                const int t = tb + 32 * u;
                if (t < lim && t != pos) vv[u] = __ldcg((const uint4 *)(vc + (size_t) t * E + h * 128 + 8 * cg));
                else vv[u] = *(const uint4 *)(v16s + 8 * cg);
            }
            #pragma unroll
            for (int u = 0; u < 4; u++) {
                const int t = tb + 32 * u;
                if (t < lim) {
                    const uint32_t w4[4] = {vv[u].x, vv[u].y, vv[u].z, vv[u].w};
                    const float p = h2f(p16[t]);
                    #pragma unroll
                    for (int e = 0; e < 8; e++) {
                        const uint16_t vh = (uint16_t)(w4[e >> 1] >> (16 * (e & 1)));
                        acc[e] = ffma(h2f(vh), p, acc[e]);
                    }
                }
            }
        }
        float4 * dst = (float4 *)(partl + (g * 8 + l) * 128 + 8 * cg);
        dst[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
        dst[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
    named_bar_sync(1, kPConsumers);
    // ---- phase 4: channel c = tid finishes with the fixed reduce tree and the double-precision tail
    if (tid < 128) {
        const int c = tid;
        float vv[8];
        #pragma unroll
        for (int l = 0; l < 8; l++) {
            const float p0 = partl[(0 * 8 + l) * 128 + c], p1 = partl[(1 * 8 + l) * 128 + c];
            const float p2 = partl[(2 * 8 + l) * 128 + c], p3 = partl[(3 * 8 + l) * 128 + c];
            vv[l] = fadd(fadd(p0, p2), fadd(p1, p3));
        }
        const float t0 = fadd(vv[0], vv[4]), t1 = fadd(vv[1], vv[5]), t2 = fadd(vv[2], vv[6]), t3 = fadd(vv[3], vv[7]);
        double sumf = (double) fadd(fadd(t0, t1), fadd(t2, t3));
        for (int t = npT; t < tcount; t++) {
            const uint16_t vh = (t == pos) ? v16s[c] : __ldcg(vc + (size_t) t * E + h * 128 + c);
            sumf += (double) fmul(h2f(vh), h2f(p16[t]));
        }
        const float ov = (float) sumf;
        a.att[h * 128 + c] = ov;
        // channels [32 w, 32 w + 32) of head h are Q8_0 block 4 h + w of the wo matmul's input
        warp_quant_block(ov, lane, a.aq_att, a.da_att, 4 * h + warp, a.dscale);
    }
}

template <int WT>
__global__ void __launch_bounds__(kPThreads, 1) k_decode_persistent(const PersistArgs a) {
    constexpr int CB = (WT == kWT_Q4_0) ? kQ4Chunk : kQ8Chunk;
    constexpr int SLOT = 16 * CB;
    extern __shared__ __align__(128) uint8_t smem[];
    const int nbq_max = a.nbqF > a.nbqE ? a.nbqF : a.nbqE;
    const PSmem lay = p_smem_layout(WT, a.NS, nbq_max, a.E, a.n_ctx);
    uint8_t * ring = smem + lay.ring;
    int * a_s = (int *)(smem + lay.a_s);
    float * da_s = (float *)(smem + lay.da_s);
    uint8_t * scratch = smem + lay.scratch;
    float * gq = (float *)(smem + lay.gq);
    double * redd = (double *)(smem + lay.red);
    float * redf = (float *)(redd + 16);
    uint64_t * full = (uint64_t *)(smem + lay.bars);
    uint64_t * empty = full + a.NS;

    const int tid = threadIdx.x, NS = a.NS;
    const int cta = blockIdx.x, n_cta = gridDim.x;
    if (tid == 0) {
        for (int i = 0; i < NS; i++) { mbar_init(&full[i], 1); mbar_init(&empty[i], kWPC); }
        mbar_fence_init();
    }
    __syncthreads();

    if (tid >= kPConsumers) {
        // ------------------------------------------------------------------ producer: one lane feeds the CTA's ring
        if (tid == kPConsumers) {
            int idx = 0;
            for (int il = 0; il < a.L; il++) {
                const PLayer & Lw = a.layers[il];
                #pragma unroll 1
                for (int ph = 0; ph < 4; ph++) {
                    const PMat & W = ph == 0 ? Lw.qkv : (ph == 1 ? Lw.wo : (ph == 2 ? Lw.w13 : Lw.w2));
                    const int n_stage = W.nbq / W.sq;
                    const uint32_t bytes = (uint32_t)(W.sq * W.TR * CB);
                    for (int round = 0; ; round++) {
                        const int n_act = p_active_groups(W.n_tiles, cta, n_cta, round);
                        if (n_act == 0) break;
                        const uint8_t * src0 = W.data + (long long)(round * kPGroups * n_cta + cta) * W.tile_bytes;
                        const long long gstride = (long long) n_cta * W.tile_bytes;      // next group's tile
                        for (int st = 0; st < n_stage; st++) {
                            for (int j = 0; j < n_act; j++, idx++) {
                                const int slot = idx % NS, use = idx / NS;
                                if (use > 0) { while (!mbar_try_wait(&empty[slot], (use - 1) & 1)) __nanosleep(64); }
                                mbar_arrive_expect_tx(&full[slot], bytes);
                                bulk_g2s(ring + (size_t) slot * SLOT, src0 + j * gstride + (size_t) st * bytes, bytes, &full[slot]);
                            }
                        }
                    }
                }
            }
        }
        return;
    }

    // ---------------------------------------------------------------------- consumers
    const int grp = tid >> 7, wig = (tid >> 5) & 3, lane = tid & 31;
    float * gq_g = gq + grp * 32;
    RingCur rc{0, 0};
    const int pos = *a.n_past;
    const int K_E = a.E, K_F = a.FF;

    for (int il = 0; il < a.L; il++) {
        const PLayer & Lw = a.layers[il];
        int * cnt = a.cnt + il * kPPhases;
        // does this CTA own tiles of a matrix?  (group 0 of CTA c owns tile c: a CTA has tiles iff cta < n_tiles)
        // ---- qkv = Wqkv . q8(rmsnorm(x) * w)
        if (cta < Lw.qkv.n_tiles) {
            if (il > 0) p_wait_counter(a.cnt + (il - 1) * kPPhases + 4, a.layers[il - 1].w2.n_tiles, tid);
            P_TRACE(0);
            p_pro_norm_quant<WT>(Lw.x_in, Lw.attn_norm, K_E, a.nb_E, a.nbqE, a_s, da_s, redd, tid);
            P_TRACE(1);
            const int done = p_run_matrix<WT, 1, EPI_STORE>(a, Lw.qkv, rc, ring, full, empty, a_s, da_s, cta, n_cta, wig, lane,
                                                            nullptr, a.qkv, 3 * a.E, gq_g, grp);
            P_TRACE(2);
            if (done) { named_bar_sync(2 + grp, kConsumers); if ((tid & 127) == 0) red_release_gpu(cnt + 0, done); }
        }
        // ---- attention, one head per CTA
        if (cta < a.H) {
            p_wait_counter(cnt + 0, Lw.qkv.n_tiles, tid);
            P_TRACE(3);
            p_attention_head(a, Lw, cta, pos, scratch, redf, redd, tid);
            named_bar_sync(1, kPConsumers);
            if (tid == 0) red_release_gpu(cnt + 1, 1);
            P_TRACE(4);
        }
        // ---- ffin = Wo . att + x
        if (cta < Lw.wo.n_tiles) {
            p_wait_counter(cnt + 1, a.H, tid);
            P_TRACE(5);
            p_pro_preq(a.aq_att, a.da_att, a.nbqE, a_s, da_s, tid);
            P_TRACE(6);
            const int done = p_run_matrix<WT, 1, EPI_RESID>(a, Lw.wo, rc, ring, full, empty, a_s, da_s, cta, n_cta, wig, lane,
                                                            Lw.x_in, a.ffin, a.E, gq_g, grp);
            P_TRACE(7);
            if (done) { named_bar_sync(2 + grp, kConsumers); if ((tid & 127) == 0) red_release_gpu(cnt + 2, done); }
        }
        // ---- gate = silu(W1 . n) * (W3 . n),  n = q8(rmsnorm(ffin) * w)
        if (cta < Lw.w13.n_tiles) {
            p_wait_counter(cnt + 2, Lw.wo.n_tiles, tid);
            P_TRACE(8);
            p_pro_norm_quant<WT>(a.ffin, Lw.ffn_norm, K_E, a.nb_E, a.nbqE, a_s, da_s, redd, tid);
            P_TRACE(9);
            const int done = p_run_matrix<WT, 2, EPI_GATEQ>(a, Lw.w13, rc, ring, full, empty, a_s, da_s, cta, n_cta, wig, lane,
                                                            nullptr, nullptr, K_F, gq_g, grp);
            P_TRACE(10);
            if (done) { named_bar_sync(2 + grp, kConsumers); if ((tid & 127) == 0) red_release_gpu(cnt + 3, done); }
        }
        // ---- x_out = W2 . gate + ffin
        if (cta < Lw.w2.n_tiles) {
            p_wait_counter(cnt + 3, Lw.w13.n_tiles, tid);
            P_TRACE(11);
            p_pro_preq(a.aq_gate, a.da_gate, a.nbqF, a_s, da_s, tid);
            P_TRACE(12);
            const int done = p_run_matrix<WT, 1, EPI_RESID>(a, Lw.w2, rc, ring, full, empty, a_s, da_s, cta, n_cta, wig, lane,
                                                            a.ffin, Lw.x_out, a.E, gq_g, grp);
            P_TRACE(13);
            if (done) { named_bar_sync(2 + grp, kConsumers); if ((tid & 127) == 0) red_release_gpu(cnt + 4, done); }
            P_TRACE(14);
        }
        if (a.trace && tid == 0) { a.trace[((size_t) blockIdx.x * a.L + il) * kPTraceSlots + 15] = (unsigned long long) rc.wait_clk; rc.wait_clk = 0; }
    }
}

}  // namespace b200
