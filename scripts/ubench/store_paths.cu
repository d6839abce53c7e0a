// Micro-benchmark: how fast can 148 SMs export [128 x 256] fp16 tiles (64 KB each) of an M x N row-major matrix, by
// which mechanism?  No MMA, no TMEM: shared memory is filled once, every "tile" is just written out.  This bounds what
// any GEMM epilogue can reach on the K = 768 ViT shapes (profiles/r1_gemm_notes.md: the epilogue, not the tensor pipe,
// bounds them).   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_paths store_paths.cu ; ./store_paths
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, const void* smem_src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"((uint64_t)tm),
                 "r"(smem_u32(smem_src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_store_1d(void* gdst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"((uint64_t)gdst), "r"(smem_u32(smem_src)),
                 "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// mode 0: TMA swizzle-128B boxes [128 rows x 64 cols], 4 per tile, one issuing thread
// mode 1: TMA swizzle-128B boxes [32 rows x 64 cols], 16 per tile, 4 issuing threads (one per warp 0..3)
// mode 2: TMA no-swizzle box [128 rows x 256 cols] (512-byte rows), 1 per tile
// mode 3: cp.async.bulk 1-D, one 512-byte row per instruction, 4 issuing threads
// mode 4: st.global.v4 from registers, one warp instruction = 4 rows x 128 B (what a 128B-swizzled staging transposes to)
// mode 5: st.global.v4 from registers, one warp instruction = one 512-byte row
// mode 6: mode 4 but the data is read back from shared memory first (ld.shared.v4 + st.global.v4)
template <int MODE>
__global__ void __launch_bounds__(256, 1) store_kernel(const __grid_constant__ CUtensorMap tm, __half* out, int M, int N) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    for (int i = threadIdx.x; i < 65536 / 16; i += blockDim.x) ((uint4*)smem)[i] = make_uint4(i, i, i, i);
    fence_proxy_async();
    __syncthreads();
    const int num_m = (M + 127) / 128, num_n = N / 256;
    const int tiles = num_m * num_n;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
        const int m0 = (t % num_m) * 128, n0 = (t / num_m) * 256;
        if (MODE == 0) {
            if (threadIdx.x == 0) {
                for (int s = 0; s < 4; ++s) { tma_store_2d(&tm, smem + s * 16384, n0 + s * 64, m0); bulk_commit(); }
                bulk_wait_read<4>();
            }
        } else if (MODE == 1) {
            if (lane == 0 && warp < 4) {
                for (int s = 0; s < 4; ++s) { tma_store_2d(&tm, smem + s * 16384 + warp * 4096, n0 + s * 64, m0 + warp * 32); bulk_commit(); }
                bulk_wait_read<4>();
            }
        } else if (MODE == 2) {
            if (threadIdx.x == 0) { tma_store_2d(&tm, smem, n0, m0); bulk_commit(); bulk_wait_read<1>(); }
        } else if (MODE == 3) {
            if (lane == 0 && warp < 4) {
                for (int r = warp * 32; r < warp * 32 + 32; ++r)
                    if (m0 + r < M) bulk_store_1d(out + (size_t)(m0 + r) * N + n0, smem + r * 512, 512);
                bulk_commit();
                bulk_wait_read<1>();
            }
        } else if (MODE == 4 || MODE == 6) {
            // 8 warps: warp w covers rows 16w..16w+15 ; per instruction 4 rows x 128 B ; 4 column slices of 64
            for (int s = 0; s < 4; ++s)
                for (int i = 0; i < 4; ++i) {
                    const int r = warp * 16 + i * 4 + (lane >> 3);
                    uint4 v = MODE == 6 ? *(const uint4*)(smem + s * 16384 + r * 128 + (((lane & 7) ^ (r & 7)) << 4))
                                        : make_uint4(t, s, i, lane);
                    if (m0 + r < M) *(uint4*)(out + (size_t)(m0 + r) * N + n0 + s * 64 + (lane & 7) * 8) = v;
                }
        } else if (MODE == 5) {
            for (int i = 0; i < 16; ++i) {
                const int r = warp * 16 + i;
                if (m0 + r < M) *(uint4*)(out + (size_t)(m0 + r) * N + n0 + lane * 8) = make_uint4(t, i, lane, 0);
            }
        }
    }
    if (MODE <= 3) bulk_wait<0>();
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static CUtensorMap make_map(EncodeTiledFn enc, void* base, int M, int N, int box_rows, int box_cols, bool swz) {
    CUtensorMap tm;
    cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)M};
    cuuint64_t strides[1] = {(cuuint64_t)N * 2};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     swz ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
    return tm;
}

template <int MODE>
static void run(const char* name, const CUtensorMap& tm, __half* out, int M, int N, int grid) {
    CK(cudaFuncSetAttribute(store_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 1024 + 1024));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) store_kernel<MODE><<<grid, 256, 66 * 1024 + 1024>>>(tm, out, M, N);
    CK(cudaDeviceSynchronize());
    const int reps = 20;
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i) store_kernel<MODE><<<grid, 256, 66 * 1024 + 1024>>>(tm, out, M, N);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double us = ms * 1e3 / reps, bytes = (double)M * N * 2;
    printf("%-58s M=%d N=%d grid=%d: %7.1f us  %6.2f TB/s  %5.1f B/clk/SM@1.9GHz\n", name, M, N, grid, us, bytes / us / 1e6,
           bytes / us / 1e6 * 1e12 / 148 / 1.9e9 / 1e0 / 1e0 * 1e-0);
}

int main() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    EncodeTiledFn enc = (EncodeTiledFn)p;
    const int shapes[3][2] = {{12000, 3072}, {12500, 2304}, {12500, 768}};
    for (int si = 0; si < 3; ++si) {
        const int M = shapes[si][0], N = shapes[si][1];
        __half* out;
        CK(cudaMalloc(&out, (size_t)M * N * 2));
        for (int grid : {148, 296}) {
            if (grid == 296 && si > 0) continue;
            CUtensorMap t128 = make_map(enc, out, M, N, 128, 64, true), t32 = make_map(enc, out, M, N, 32, 64, true),
                        tbig = make_map(enc, out, M, N, 128, 256, false);
            run<0>("0 TMA sw128 box 128x64 (4/tile, 1 thread)", t128, out, M, N, grid);
            run<1>("1 TMA sw128 box 32x64 (16/tile, 4 threads)", t32, out, M, N, grid);
            run<2>("2 TMA no-swizzle box 128x256 (1/tile)", tbig, out, M, N, grid);
            run<3>("3 cp.async.bulk 1-D 512 B rows (128/tile, 4 threads)", t128, out, M, N, grid);
            run<4>("4 st.global.v4, 4 rows x 128 B per warp instr (regs)", t128, out, M, N, grid);
            run<5>("5 st.global.v4, 1 row x 512 B per warp instr (regs)", t128, out, M, N, grid);
            run<6>("6 ld.shared.v4 + st.global.v4, 4 rows x 128 B", t128, out, M, N, grid);
        }
        CK(cudaFree(out));
    }
    return 0;
}
