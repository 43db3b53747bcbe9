// fastgemm.cuh -- K2: prefill weight matmul on the 5th-generation tensor cores ("fast mode").
//
//   Y[token][row] = sum_k  W[row][k] * X[token][k]        W: Q4_0 (packed layout of kernels.cuh), X: activations
//
// tcgen05.mma (kind::f16, M=128, N=128, K=16, cta_group::1) with the fp32 accumulator in TENSOR MEMORY:
//   * the producer lane streams the packed Q4_0 chunks of four 32-row tiles (= one 128-row M tile) with 1-D TMA bulk
//     copies into a raw ring -- the same 18 B/block HBM traffic as the decode kernel, nothing is dequantised in HBM;
//   * four dequant warps expand each 4-block quad to fp16 IN SHARED MEMORY, writing the K-major SWIZZLE_128B layout the
//     UMMA shared-memory descriptor expects (fp16 magic-number nibble conversion, one HMUL2 by the block scale), and
//     copy the matching activation tile; `fence.proxy.async` hands the tiles to the tensor core's async proxy;
//   * ONE elected thread issues 8 tcgen05.mma per 128-wide K block and `tcgen05.commit`s to an mbarrier that recycles
//     the stage; after the last K block the four warps read the accumulator back with tcgen05.ld (32 lanes x 16
//     columns per instruction) and apply the fused epilogue (store | +residual | SiLU-gate).
// Numerics ("fast mode", tolerance-checked, NOT bit-exact): activations go through the reference's Q8_0
// quantisation (k_prep_q8_f16) and both operands are rounded to fp16 (<= 2^-11 relative each); products are exact
// in fp32 and accumulated in fp32 in hardware order.  Measured: 2.7e-4 relative RMS on one matmul; ~5e-3 per layer on
// hidden states, dominated by Q8_0 codes of the NEXT matmul's input flipping by one step (tests/test_gpu_fast_prefill.py).
#pragma once
#include "kernels.cuh"

namespace b200 {

constexpr int kFgM = 128, kFgN = 128, kFgK = 128;     // CTA tile: 128 weight rows x 128 tokens, K block of 128 (one quad)
constexpr int kFgStages = 2;
constexpr int kFgRawBytes = 4 * 4 * kQ4Chunk;         // 4 tiles x 4 row-groups x 576 B = one quad of 128 rows
constexpr int kFgABytes = kFgM * kFgK * 2;            // 32 KB: two [128 x 64] K-major SW128 sub-tiles
constexpr int kFgBBytes = kFgN * kFgK * 2;            // 32 KB
constexpr int kFgStageBytes = kFgRawBytes + kFgABytes + kFgBBytes;
constexpr int kFgSmem = kFgStages * kFgStageBytes + 1024 /*align*/ + 256;

// ---- activation pre-pass: [RMSNorm * w ->] Q8_0 quantise -> dequantise -> fp16 row (one warp per 32-block) --------
struct PrepArgs { const float * x; int ldx; const float * norm_w; uint16_t * xh; int K, N; };

template <bool NORM>
__global__ void __launch_bounds__(256) k_prep_q8_f16(const PrepArgs a) {
    __shared__ double red[8];
    grid_dep_wait();
    const int n = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float * x = a.x + (size_t) n * a.ldx;
    float scale = 1.f;
    if (NORM) {
        double s = 0.0;
        for (int i = tid; i < a.K; i += 256) s += (double) fmul(x[i], x[i]);
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) red[warp] = s;
        __syncthreads();
        double tot = 0.0;
        for (int i = 0; i < 8; i++) tot += red[i];
        scale = __fdiv_rn(1.0f, __fsqrt_rn(fadd((float)(tot / (double) a.K), 1e-6f)));
    }
    for (int b = warp; b < a.K / 32; b += 8) {
        float v = x[b * 32 + lane];
        if (NORM) v = fmul(fmul(v, scale), a.norm_w[b * 32 + lane]);
        float amax = fabsf(v);
        for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
        const float d = h2f(f2h(__fdiv_rn(amax, 127.f)));
        const float id = amax != 0.f ? __fdiv_rn(127.f, amax) : 0.f;
        const float q = (float) rint_small(fmul(v, id));
        a.xh[(size_t) n * a.K + b * 32 + lane] = f2h(fmul(q, d));
    }
}

// ---- tcgen05 helpers -----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
    // K-major, SWIZZLE_128B, 8-row groups 1024 B apart (cute::UMMA::SmemDescriptor: start>>4 | LBO 1 | SBO 64 | version 1 | layout 2)
    return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t) 1 << 16) | ((uint64_t) 64 << 32) | ((uint64_t) 1 << 46) | ((uint64_t) 2 << 61);
}
__device__ __forceinline__ uint32_t umma_idesc_f16_f32(int M, int N) {
    // c_format F32 (1) at bit 4, a/b format F16 (0), K-major both, n_dim = N>>3 at bit 17, m_dim = M>>4 at bit 24
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p; }"
                 :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t * bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after()  { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

enum { FG_STORE = 0, FG_RESID = 1, FG_GATE = 2 };

struct FastGemmArgs {
    PackedW W;                  // G=1 packing for STORE/RESID (TR = 4), G=2 packing (TR = 8) for GATE
    const uint16_t * xh;        // [N][K] fp16 (k_prep_q8_f16)
    const float * resid; int ldr;
    float * y; int ldy;
    int N, out_rows;
    const uint16_t * tsilu;
};

template <int EPI>
__global__ void __launch_bounds__(160, 1) k_gemm_q4_tc(const FastGemmArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t * smem = (uint8_t *)(((uintptr_t) smem_raw + 1023) & ~(uintptr_t) 1023);       // SWIZZLE_128B tiles need 1 KB alignment
    uint64_t * bars = (uint64_t *)(smem + kFgStages * kFgStageBytes);
    uint64_t * raw_full = bars, * ab_full = bars + 2, * stage_free = bars + 4, * acc_full = bars + 6;
    uint32_t * tmem_slot = (uint32_t *)(bars + 8);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int mt = blockIdx.x, nt = blockIdx.y;             // 128-row tile, 128-token tile
    const int nbq = a.W.nbq, K = a.W.K;
    const int TRp = a.W.TR;                                 // row-groups per packed tile (4 or 8)
    const int tiles_per_m = 16 / TRp;                       // packed tiles per 128 rows (16 row-groups)

    if (tid == 0) {
        for (int s = 0; s < kFgStages; s++) { mbar_init(&raw_full[s], 1); mbar_init(&ab_full[s], 4); mbar_init(&stage_free[s], 1); }
        mbar_init(acc_full, 1);
        mbar_fence_init();
    }
    if (warp == 4) {                                        // TMEM: 128 columns of fp32 accumulator
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" :: "r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 4) {
        if (lane == 0) {
            // ---------------------------------------------------------------- TMA producer + MMA issuer (one thread)
            grid_dep_launch();
            const uint32_t idesc = umma_idesc_f16_f32(kFgM, kFgN);
            for (int kb = 0; kb < nbq + 1; kb++) {
                if (kb < nbq) {                              // stream the raw quad of 128 rows for K block kb
                    const int s = kb % kFgStages, use = kb / kFgStages;
                    if (use > 0) mbar_wait(&stage_free[s], (use - 1) & 1);
                    uint8_t * raw = smem + (size_t) s * kFgStageBytes;
                    const uint32_t per_tile = (uint32_t) TRp * kQ4Chunk;
                    mbar_arrive_expect_tx(&raw_full[s], (uint32_t) kFgRawBytes);
                    for (int t = 0; t < tiles_per_m; t++) {
                        const uint8_t * src = a.W.data + (long long)(mt * tiles_per_m + t) * a.W.tile_bytes + (size_t) kb * per_tile;
                        bulk_g2s(raw + (size_t) t * per_tile, src, per_tile, &raw_full[s]);
                    }
                }
                if (kb > 0) {                                // issue the MMAs of K block kb-1
                    const int j = kb - 1, s = j % kFgStages;
                    mbar_wait(&ab_full[s], (j / kFgStages) & 1);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(smem + (size_t) s * kFgStageBytes + kFgRawBytes);
                    const uint32_t b_addr = a_addr + kFgABytes;
                    #pragma unroll
                    for (int k16 = 0; k16 < 8; k16++) {
                        const uint32_t sub = (k16 >> 2) * (kFgM * 128), off = (k16 & 3) * 32;      // sub-tile of 64 K, 32 B per UMMA_K
                        umma_f16(tmem_base, umma_desc_k_sw128(a_addr + sub + off), umma_desc_k_sw128(b_addr + sub + off),
                                 idesc, (j > 0 || k16 > 0) ? 1u : 0u);
                    }
                    umma_commit(&stage_free[s]);
                    if (j == nbq - 1) umma_commit(acc_full);
                }
            }
        }
    } else {
        // -------------------------------------------------------------------- dequant warps (128 threads)
        grid_dep_wait();
        const int r8 = lane >> 2, w = lane & 3;
        for (int kb = 0; kb < nbq; kb++) {
            const int s = kb % kFgStages, use = kb / kFgStages;
            uint8_t * stage = smem + (size_t) s * kFgStageBytes;
            uint8_t * A = stage + kFgRawBytes, * B = A + kFgABytes;
            // B tile: tokens nt*128 .. +127, K block kb: [128 tokens][128 halfs] -> two SW128 sub-tiles (stage is free: the
            // producer waited on stage_free before re-arming raw_full, and we wait on raw_full below before touching A)
            mbar_wait(&raw_full[s], use & 1);
            for (int c = tid; c < kFgN * 16; c += 128) {     // 16-byte chunks: 128 rows x 16 chunks
                const int row = c >> 4, ch = c & 15, tok = nt * kFgN + row;
                uint4 v = make_uint4(0, 0, 0, 0);
                if (tok < a.N && kb * kFgK + ch * 8 < K) v = *(const uint4 *)(a.xh + (size_t) tok * K + kb * kFgK + ch * 8);
                const int sub = ch >> 3, c8 = ch & 7;
                *(uint4 *)(B + sub * (kFgN * 128) + row * 128 + ((c8 ^ (row & 7)) << 4)) = v;
            }
            // A tile: this warp expands row-groups 4*warp .. 4*warp+3 (rows 32*warp .. +31) of the quad
            #pragma unroll
            for (int gi = 0; gi < 4; gi++) {
                const int g = warp * 4 + gi;                 // row-group 0..15 of the M tile, in packed order
                const uint8_t * ch = stage + (size_t) g * kQ4Chunk;
                const uint4 wv = *(const uint4 *)(ch + lane * 16);
                const uint2 sc = *(const uint2 *)(ch + 512 + r8 * 8);
                const uint32_t ww[4] = {wv.x, wv.y, wv.z, wv.w};
                const uint32_t sw[2] = {sc.x, sc.y};
                const int row = g * 8 + r8;                  // row within the 128-row tile (packed order)
                #pragma unroll
                for (int bq = 0; bq < 4; bq++) {
                    const uint32_t x = ww[bq] ^ 0x88888888u;             // back to offset-binary nibbles n = v + 8
                    const uint32_t dh = (sw[bq >> 1] >> (16 * (bq & 1))) & 0xFFFFu;
                    const __half2 d2 = __halves2half2(__ushort_as_half((unsigned short) dh), __ushort_as_half((unsigned short) dh));
                    const __half2 off = __halves2half2(__ushort_as_half((unsigned short) 0x6408), __ushort_as_half((unsigned short) 0x6408));   // 1032.0
                    uint32_t h[4];                            // {e0,e2} {e16,e18} {e1,e3} {e17,e19} relative to 4w
                    #pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const uint32_t m = ((x >> (4 * i)) & 0x000F000Fu) | 0x64006400u;      // 1024 + n, exact in fp16
                        const __half2 v = __hmul2(__hsub2(*(const __half2 *) &m, off), d2);     // (n - 8) * d, one rounding
                        h[i] = *(const uint32_t *) &v;
                    }
                    // reorder pairs to consecutive k: lo = e0 e1 e2 e3, hi = e16 e17 e18 e19
                    const uint32_t lo0 = __byte_perm(h[0], h[2], 0x5410), lo1 = __byte_perm(h[0], h[2], 0x7632);
                    const uint32_t hi0 = __byte_perm(h[1], h[3], 0x5410), hi1 = __byte_perm(h[1], h[3], 0x7632);
                    const int klo = bq * 32 + 4 * w, khi = klo + 16;                      // k within the 128-wide block
                    {
                        const int sub = klo >> 6, kk = klo & 63, c8 = kk >> 3, within = (kk & 7) * 2;
                        *(uint2 *)(A + sub * (kFgM * 128) + row * 128 + ((c8 ^ (row & 7)) << 4) + within) = make_uint2(lo0, lo1);
                    }
                    {
                        const int sub = khi >> 6, kk = khi & 63, c8 = kk >> 3, within = (kk & 7) * 2;
                        *(uint2 *)(A + sub * (kFgM * 128) + row * 128 + ((c8 ^ (row & 7)) << 4) + within) = make_uint2(hi0, hi1);
                    }
                }
            }
            fence_proxy_async();                             // generic-proxy writes -> visible to the tensor core's async proxy
            __syncwarp();
            if (lane == 0) mbar_arrive(&ab_full[s]);
        }
        // -------------------------------------------------------------------- epilogue: TMEM -> registers -> global
        mbar_wait(acc_full, 0);
        tc_fence_after();
        const int m = warp * 32 + lane;                      // accumulator lane = row within the M tile (packed order)
        #pragma unroll 1
        for (int c0 = 0; c0 < kFgN; c0 += 16) {
            uint32_t v[16];
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t) c0;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                           "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                         : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            #pragma unroll
            for (int j = 0; j < 16; j++) {
                const int tok = nt * kFgN + c0 + j;
                float val = __uint_as_float(v[j]);
                if (EPI == FG_GATE) {
                    // packed G=2 order: row-groups alternate w1 / w3, so lane l (w1) pairs with lane l^8 (w3) of the same row
                    const float other = __shfl_xor_sync(0xffffffffu, val, 8);
                    const int row = (mt * 8 + (m >> 4)) * 8 + (m & 7);
                    if (!(m & 8) && tok < a.N && row < a.out_rows)
                        a.y[(size_t) tok * a.ldy + row] = fmul(h2f(a.tsilu[f2h(val)]), other);
                } else {
                    const int row = mt * kFgM + m;
                    if (tok < a.N && row < a.out_rows) {
                        if (EPI == FG_RESID) val = fadd(val, a.resid[(size_t) tok * a.ldr + row]);
                        a.y[(size_t) tok * a.ldy + row] = val;
                    }
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" :: "r"(tmem_base) : "memory");
}

}  // namespace b200
