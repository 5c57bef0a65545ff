// Probe: which cp.async.bulk.tensor variant works for a uint8 [frame][row][pitch] tensor.
// usage: tma_probe rank c0 bw bh [frameStridePad]
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int RANK>
__global__ void probe(const __grid_constant__ CUtensorMap pmap, int c0, int c1, int c2, int bytes, unsigned char* out) {
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ __align__(8) unsigned long long bar;
    const unsigned b = smem_u32(&bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
        if (RANK == 3)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(sm)),
                         "l"(&pmap), "r"(c0), "r"(c1), "r"(c2), "r"(b) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(sm)),
                         "l"(&pmap), "r"(c0), "r"(c1), "r"(b) : "memory");
    }
    asm volatile(
        "{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(b)
        : "memory");
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = sm[i];
}

int main(int argc, char** argv) {
    const int rank = atoi(argv[1]), c0 = atoi(argv[2]), bw = atoi(argv[3]), bh = atoi(argv[4]);
    const int pitch = 816, rows = 518, frames = 2;
    size_t frameBytes = (size_t)pitch * rows;
    if (argc > 5 && atoi(argv[5])) frameBytes += 256 - frameBytes % 256;
    std::vector<unsigned char> h(frameBytes * frames);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (unsigned char)((i * 2654435761u) >> 13);
    unsigned char *d, *out;
    cudaMalloc(&d, h.size());
    cudaMalloc(&out, bw * bh);
    cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    EncodeFn encode = (EncodeFn)fn;
    alignas(64) CUtensorMap map;
    const cuuint64_t dims[3] = {(cuuint64_t)pitch, (cuuint64_t)rows * (rank == 2 ? frames : 1), (cuuint64_t)frames};
    const cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)frameBytes};
    const cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, rank, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("rank %d c0 %d box %dx%d frameBytes %zu: encode -> %d; ", rank, c0, bw, bh, frameBytes, (int)r);
    std::vector<unsigned char> got(bw * bh);
    const int c1 = 16, c2 = rank == 3 ? 1 : 0;
    cudaMemset(out, 0xee, bw * bh);
    if (rank == 3) probe<3><<<1, 128, bw * bh + 128>>>(map, c0, c1, c2, bw * bh, out);
    else probe<2><<<1, 128, bw * bh + 128>>>(map, c0, c1, c2, bw * bh, out);
    cudaError_t e = cudaDeviceSynchronize();
    printf("%s; ", cudaGetErrorString(e));
    if (e != cudaSuccess) { printf("\n"); return 1; }
    cudaMemcpy(got.data(), out, bw * bh, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int y = 0; y < bh; ++y)
        for (int x = 0; x < bw; ++x)
            bad += got[y * bw + x] != h[(size_t)c2 * frameBytes + (size_t)(c1 + y) * pitch + c0 + x];
    printf("%d mismatching bytes\n", bad);
    return 0;
}
