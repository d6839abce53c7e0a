// Persistent warp-specialised tcgen05 GEMM for sm_100a.
//
//   D[M,N] = epilogue( A[M,K] . B[N,K]^T )     A, B fp16 K-major; accumulate fp32 in TMEM
//
// One CTA per SM, 256 threads:
//   warp 0   TMA producer      (one lane): global -> 128B-swizzled smem ring, STAGES deep
//   warp 1   MMA issuer        (one lane): tcgen05.mma 128 x BN x 16, accumulators double-buffered in TMEM
//   warp 2   TMEM allocator
//   warps 4-7 epilogue         : tcgen05.ld -> scale/bias/activation/addend/residual -> 16-byte global stores
// Three mbarrier pipelines: smem full/empty (TMA <-> MMA), TMEM full/empty (MMA <-> epilogue); tiles are
// walked m-fastest so that co-resident CTAs share the same weight (B) tile through L2.
//
// Used for every dense contraction of the hot path: the ViT patch-embed / QKV / out-proj / FFN GEMMs
// (reference: third-party clip `VisionTransformer.forward`, called at models/CLIP/extract_clip.py:128).
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "internal.h"

namespace vf {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;   // 64 fp16 = one 128-byte swizzle row

template <int BN, int STAGES>
struct GemmCfg {
    static constexpr uint32_t A_BYTES = BM * BK * 2;
    static constexpr uint32_t B_BYTES = BN * BK * 2;
    static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr uint32_t TMEM_COLS = 2 * BN;   // two accumulator stages; power of two in [32,512]
    static constexpr uint32_t BAR_BYTES = (2 * STAGES + 4) * 8 + 16;
    static constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + 1024;   // + alignment slack
    static_assert(TMEM_COLS == 64 || TMEM_COLS == 128 || TMEM_COLS == 256 || TMEM_COLS == 512, "TMEM columns");
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == VF_ACT_QUICKGELU) {
        // x * sigmoid(1.702 x)
        return __fdividef(v, 1.0f + __expf(-1.702f * v));
    } else if (act == VF_ACT_RELU) {
        return fmaxf(v, 0.0f);
    }
    return v;
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(256, 1)
gemm_f16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmEpi ep,
                const int M, const int N, const int K) {
    using Cfg = GemmCfg<BN, STAGES>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * Cfg::A_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* tfull = empty + STAGES;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_m = (M + BM - 1) / BM;
    const int num_n = (N + BN - 1) / BN;
    const int num_tiles = num_m * num_n;
    const int num_k = (K + BK - 1) / BK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full[i], 1);
            mbar_init(&empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull[i], 1);
            mbar_init(&tempty[i], 4);   // one arrival per epilogue warp
        }
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m_blk = tile % num_m, n_blk = tile / num_m;
                for (int kb = 0; kb < num_k; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    mbar_expect_tx(&full[stage], Cfg::STAGE_BYTES);
                    tma_load_2d(sA + stage * Cfg::A_BYTES, &tmA, &full[stage], kb * BK, m_blk * BM);
                    tma_load_2d(sB + stage * Cfg::B_BYTES, &tmB, &full[stage], kb * BK, n_blk * BN);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_f16(BM, BN, 0);
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                mbar_wait(&tempty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = 0; kb < num_k; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint64_t adesc = umma_desc_sw128(sA + stage * Cfg::A_BYTES);
                    const uint64_t bdesc = umma_desc_sw128(sB + stage * Cfg::B_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        // +32 bytes (encoded >>4 = 2) per 16-element K step inside the 128-byte swizzle row
                        umma_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(&empty[stage]);   // frees the smem slot once these MMAs have read it
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tfull[acc]);          // accumulator complete -> epilogue
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------ epilogue (128 threads = 128 TMEM lanes)
        const int ew = warp & 3;   // TMEM lane quarter this warp may access
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int m_blk = tile % num_m, n_blk = tile / num_m;
            mbar_wait(&tfull[acc], acc_phase);
            tc_fence_after();
            const int m = m_blk * BM + ew * 32 + lane;
            const bool row_ok = m < M;
            int64_t orow = m;
            int arow = 0;
            if (ep.gin > 0) {
                const int g = m / ep.gin, r = m - g * ep.gin;
                orow = int64_t(g) * ep.gout + ep.goff + r;
                arow = ep.goff + r;
            }
            const uint32_t t_row = tmem_base + acc * BN + (uint32_t(ew * 32) << 16);
#pragma unroll 1
            for (int c = 0; c < BN; c += 32) {
                uint32_t raw[32];
                tmem_ld_32x32(t_row + c, raw);
                tmem_ld_wait();
                const int n0 = n_blk * BN + c;
                if (row_ok && n0 < N) {
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
                    if (ep.scale) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 s = __ldg(reinterpret_cast<const float4*>(ep.scale + n0 + j));
                            v[j] *= s.x; v[j + 1] *= s.y; v[j + 2] *= s.z; v[j + 3] *= s.w;
                        }
                    }
                    if (ep.bias) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias + n0 + j));
                            v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
                        }
                    }
                    if (ep.act != VF_ACT_NONE) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = apply_act(v[j], ep.act);
                    }
                    if (ep.addend) {
                        const float* ap = ep.addend + int64_t(arow) * N + n0;
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 a = __ldg(reinterpret_cast<const float4*>(ap + j));
                            v[j] += a.x; v[j + 1] += a.y; v[j + 2] += a.z; v[j + 3] += a.w;
                        }
                    }
                    if (ep.residual) {
                        const float* rp = ep.residual + orow * ep.ldr + n0;
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 r = *reinterpret_cast<const float4*>(rp + j);
                            v[j] += r.x; v[j + 1] += r.y; v[j + 2] += r.z; v[j + 3] += r.w;
                        }
                    }
                    if (ep.out_f32) {
                        float* op = reinterpret_cast<float*>(ep.out) + orow * ep.ldo + n0;
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            *reinterpret_cast<float4*>(op + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    } else {
                        __half* op = reinterpret_cast<__half*>(ep.out) + orow * ep.ldo + n0;
#pragma unroll
                        for (int j = 0; j < 32; j += 8) {
                            uint4 pk;
                            pk.x = pack_half2(v[j], v[j + 1]);
                            pk.y = pack_half2(v[j + 2], v[j + 3]);
                            pk.z = pack_half2(v[j + 4], v[j + 5]);
                            pk.w = pack_half2(v[j + 6], v[j + 7]);
                            *reinterpret_cast<uint4*>(op + j) = pk;
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}


// =====================================================================================================
// v2: CTA pairs (tcgen05 cta_group::2).  A cluster of two CTAs (one TPC) owns a 256 x BN output tile: each CTA stages
// its own 128 rows of A and HALF of the B tile (BN/2 rows), the leader issues one UMMA of M=256 that reads both
// halves, and each CTA's TMEM receives the 128 x BN block of its rows.  Versus the single-CTA kernel this halves the
// B bytes every SM pulls from L2 (the measured limiter of v1: ~41 B/clk/SM of L2->SM traffic at 47 % tensor-pipe
// utilisation) and frees shared memory for a staged, fully coalesced epilogue:
//   warps 4-11 (8 epilogue warps): tcgen05.ld (thread = row) -> padded smem transpose -> (lane = 4 columns) ->
//   scale / bias / activation / addend / residual -> 128-byte-per-row global stores.
template <int BN, int STAGES>
struct Gemm2Cfg {
    static constexpr uint32_t A_BYTES = BM * BK * 2;          // 128 rows of A per CTA
    static constexpr uint32_t B_BYTES = (BN / 2) * BK * 2;    // half of the B tile per CTA
    static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr uint32_t TMEM_COLS = 2 * BN;
    static constexpr uint32_t EPI_WARPS = 8;
    static constexpr uint32_t STG_LD = 36;                    // words per staged row (32 + 4 pad: conflict-free)
    static constexpr uint32_t STG_BYTES = EPI_WARPS * 32 * STG_LD * 4;
    static constexpr uint32_t BAR_BYTES = (2 * STAGES + 4) * 8 + 16;
    static constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + STG_BYTES + BAR_BYTES + 1024;
    static_assert(TMEM_COLS == 128 || TMEM_COLS == 256 || TMEM_COLS == 512, "TMEM columns");
    static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

template <int BN, int STAGES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(384, 1)
gemm_f16_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const GemmEpi ep, const int M, const int N, const int K) {
    using Cfg = Gemm2Cfg<BN, STAGES>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * Cfg::A_BYTES;
    float* stg = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES + Cfg::STG_BYTES);
    uint64_t* empty = full + STAGES;
    uint64_t* tfull = empty + STAGES;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t cta = cluster_ctarank();        // 0 = leader (issues the MMAs), 1 = peer
    const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
    const int num_m = (M + 2 * BM - 1) / (2 * BM);
    const int num_n = (N + BN - 1) / BN;
    const int num_tiles = num_m * num_n;
    const int num_k = (K + BK - 1) / BK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full[i], 1);      // leader's own arrive.expect_tx; bytes of BOTH CTAs are credited here
            mbar_init(&empty[i], 1);     // multicast tcgen05.commit from the leader
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull[i], 1);                       // multicast tcgen05.commit
            mbar_init(&tempty[i], 2 * Cfg::EPI_WARPS);     // (leader only is waited on) every epilogue warp of both CTAs
        }
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc_2sm(tmem_slot, Cfg::TMEM_COLS);
    tc_fence_before();
    cluster_sync_all();      // barriers of both CTAs initialised before any remote arrive / TMA credit
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer (both CTAs)
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = pair; tile < num_tiles; tile += num_pairs) {
                const int m_blk = tile % num_m, n_blk = tile / num_m;
                const int m0 = m_blk * 2 * BM + int(cta) * BM;
                const int n0 = n_blk * BN + int(cta) * (BN / 2);
                for (int kb = 0; kb < num_k; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    if (cta == 0) mbar_expect_tx(&full[stage], 2 * Cfg::STAGE_BYTES);
                    const uint32_t bar = mapa_u32(smem_u32(&full[stage]), 0);
                    tma_load_2d_2sm(sA + stage * Cfg::A_BYTES, &tmA, bar, kb * BK, m0);
                    tma_load_2d_2sm(sB + stage * Cfg::B_BYTES, &tmB, bar, kb * BK, n0);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer (leader CTA, one lane)
        if (cta == 0 && lane == 0) {
            constexpr uint32_t idesc = umma_idesc_f16(2 * BM, BN, 0);
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
            for (int tile = pair; tile < num_tiles; tile += num_pairs) {
                mbar_wait(&tempty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = 0; kb < num_k; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint64_t adesc = umma_desc_sw128(sA + stage * Cfg::A_BYTES);
                    const uint64_t bdesc = umma_desc_sw128(sB + stage * Cfg::B_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)
                        umma_f16_2sm(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                    umma_commit_2sm(&empty[stage], 3);    // free this smem slot in both CTAs
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit_2sm(&tfull[acc], 3);           // accumulators of both CTAs complete
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------ epilogue: 8 warps per CTA
        const int e = warp - 4;
        const int q = e & 3;                   // TMEM lane quarter (warp id % 4)
        const int half = e >> 2;               // which half of the BN columns this warp drains
        float* my = stg + e * 32 * Cfg::STG_LD;
        const int rq = lane >> 3, cq = (lane & 7) * 4;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = pair; tile < num_tiles; tile += num_pairs) {
            const int m_blk = tile % num_m, n_blk = tile / num_m;
            const int mrow0 = m_blk * 2 * BM + int(cta) * BM + q * 32;
            // the 8 rows this lane stores (after the transpose): mrow0 + i*4 + rq
            int64_t orow[8];
            int arow[8];
            bool rok[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int m = mrow0 + i * 4 + rq;
                rok[i] = m < M;
                orow[i] = m;
                arow[i] = 0;
                if (ep.gin > 0) {
                    const int g = m / ep.gin, r = m - g * ep.gin;
                    orow[i] = int64_t(g) * ep.gout + ep.goff + r;
                    arow[i] = ep.goff + r;
                }
            }
            mbar_wait(&tfull[acc], acc_phase);
            tc_fence_after();
            const uint32_t t_row = tmem_base + acc * BN + (uint32_t(q * 32) << 16);
#pragma unroll 1
            for (int c = half * (BN / 2); c < (half + 1) * (BN / 2); c += 32) {
                uint32_t raw[32];
                tmem_ld_32x32(t_row + c, raw);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<uint4*>(my + lane * Cfg::STG_LD + j) =
                        make_uint4(raw[j], raw[j + 1], raw[j + 2], raw[j + 3]);
                __syncwarp();
                const int n = n_blk * BN + c + cq;       // first of this lane's 4 columns
                if (n < N) {
                    float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), bi = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (ep.scale) sc = __ldg(reinterpret_cast<const float4*>(ep.scale + n));
                    if (ep.bias) bi = __ldg(reinterpret_cast<const float4*>(ep.bias + n));
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        if (!rok[i]) continue;
                        float4 v = *reinterpret_cast<const float4*>(my + (i * 4 + rq) * Cfg::STG_LD + cq);
                        v.x = fmaf(v.x, sc.x, bi.x); v.y = fmaf(v.y, sc.y, bi.y);
                        v.z = fmaf(v.z, sc.z, bi.z); v.w = fmaf(v.w, sc.w, bi.w);
                        if (ep.act != VF_ACT_NONE) {
                            v.x = apply_act(v.x, ep.act); v.y = apply_act(v.y, ep.act);
                            v.z = apply_act(v.z, ep.act); v.w = apply_act(v.w, ep.act);
                        }
                        if (ep.addend) {
                            const float4 a = __ldg(reinterpret_cast<const float4*>(ep.addend + int64_t(arow[i]) * N + n));
                            v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
                        }
                        if (ep.residual) {
                            const float4 r = *reinterpret_cast<const float4*>(ep.residual + orow[i] * ep.ldr + n);
                            v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w;
                        }
                        if (ep.out_f32) {
                            *reinterpret_cast<float4*>(reinterpret_cast<float*>(ep.out) + orow[i] * ep.ldo + n) = v;
                        } else {
                            *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(ep.out) + orow[i] * ep.ldo + n) =
                                make_uint2(pack_half2(v.x, v.y), pack_half2(v.z, v.w));
                        }
                    }
                }
                __syncwarp();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_remote(&tempty[acc], 0);   // tell the leader's MMA issuer this stage is drained
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    }

    tc_fence_before();
    cluster_sync_all();      // the peer's smem / barriers stay alive until the leader's MMAs and commits are done
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
    return fn;
}

template <int BN, int STAGES>
int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmEpi& ep, int M, int N, int K,
                cudaStream_t stream) {
    using Cfg = GemmCfg<BN, STAGES>;
    static bool attr_set = false;
    if (!attr_set) {
        VF_CUDA(cudaFuncSetAttribute(gemm_f16_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     Cfg::SMEM_BYTES));
        attr_set = true;
    }
    const int tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
    const int sms = device_sm_count();
    const int grid = tiles < sms ? tiles : sms;
    gemm_f16_kernel<BN, STAGES><<<grid, 256, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, ep, M, N, K);
    VF_CUDA(cudaGetLastError());
    return VF_OK;
}

template <int BN, int STAGES>
int launch_gemm_pair(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmEpi& ep, int M, int N, int K,
                     cudaStream_t stream) {
    using Cfg = Gemm2Cfg<BN, STAGES>;
    static bool attr_set[64] = {false};
    int dev = 0;
    VF_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        VF_CUDA(cudaFuncSetAttribute(gemm_f16_pair_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     Cfg::SMEM_BYTES));
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    const int tiles = ((M + 2 * BM - 1) / (2 * BM)) * ((N + BN - 1) / BN);
    const int pairs = device_sm_count() / 2;
    const int grid = 2 * (tiles < pairs ? tiles : pairs);
    gemm_f16_pair_kernel<BN, STAGES><<<grid, 384, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, ep, M, N, K);
    VF_CUDA(cudaGetLastError());
    return VF_OK;
}

// VF_GEMM=1cta selects the single-CTA kernel (kept for A/B measurements); default is the CTA-pair kernel
bool use_single_cta() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("VF_GEMM");
        v = (e && strcmp(e, "1cta") == 0) ? 1 : 0;
    }
    return v == 1;
}

}  // namespace

int device_sm_count() {
    static int sms[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (sms[dev] == 0) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        sms[dev] = v;
    }
    return sms[dev];
}

int make_tmap_2d_f16(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t row_pitch_bytes,
                     uint32_t box_rows, uint32_t box_cols) {
    EncodeTiledFn enc = get_encode_tiled();
    if (!enc) return fail(VF_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
    if ((reinterpret_cast<uintptr_t>(base) & 15) || (row_pitch_bytes & 15))
        return fail(VF_ERR_INVALID, "TMA operand must be 16-byte aligned with a 16-byte multiple row pitch");
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {row_pitch_bytes};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(VF_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", int(r));
    return VF_OK;
}

int gemm_f16(const __half* A, int lda, const __half* B, int ldb, int M, int N, int K, const GemmEpi& ep,
             cudaStream_t stream) {
    if (M <= 0 || N <= 0 || K <= 0) return fail(VF_ERR_INVALID, "gemm: empty problem %dx%dx%d", M, N, K);
    if (N % 8) return fail(VF_ERR_INVALID, "gemm: N=%d must be a multiple of 8", N);
    if (K % 8 || lda % 8 || ldb % 8) return fail(VF_ERR_INVALID, "gemm: K/lda/ldb must be multiples of 8");
    if (ep.out_f32 ? (ep.ldo % 4) : (ep.ldo % 8)) return fail(VF_ERR_INVALID, "gemm: ldo breaks 16-byte stores");
    if (ep.residual && (ep.ldr % 4)) return fail(VF_ERR_INVALID, "gemm: ldr must be a multiple of 4");
    // tile-N choice: widest tile that divides N (all ViT widths are multiples of 256)
    const int bn = (N % 256 == 0) ? 256 : (N % 128 == 0) ? 128 : 64;
    CUtensorMap tmA, tmB;
    VF_TRY(make_tmap_2d_f16(&tmA, A, uint64_t(M), uint64_t(K), uint64_t(lda) * 2, BM, BK));
    if (!use_single_cta()) {
        const int bn2 = (N > 128) ? 256 : (N > 64) ? 128 : 64;     // pair-tile width; B box = half of it
        VF_TRY(make_tmap_2d_f16(&tmB, B, uint64_t(N), uint64_t(K), uint64_t(ldb) * 2, uint32_t(bn2 / 2), BK));
        if (bn2 == 256) return launch_gemm_pair<256, 5>(tmA, tmB, ep, M, N, K, stream);
        if (bn2 == 128) return launch_gemm_pair<128, 6>(tmA, tmB, ep, M, N, K, stream);
        return launch_gemm_pair<64, 8>(tmA, tmB, ep, M, N, K, stream);
    }
    if (N % 32) return fail(VF_ERR_INVALID, "gemm(1cta): N=%d must be a multiple of 32", N);
    VF_TRY(make_tmap_2d_f16(&tmB, B, uint64_t(N), uint64_t(K), uint64_t(ldb) * 2, uint32_t(bn), BK));
    if (bn == 256) return launch_gemm<256, 4>(tmA, tmB, ep, M, N, K, stream);
    if (bn == 128) return launch_gemm<128, 6>(tmA, tmB, ep, M, N, K, stream);
    return launch_gemm<64, 8>(tmA, tmB, ep, M, N, K, stream);
}

}  // namespace vf
