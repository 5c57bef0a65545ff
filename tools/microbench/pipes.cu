// Issue-rate microbenchmark for the integer/packed ops the FAST kernel is made of (B200, sm_100a).
// Each kernel runs ITER x 8 independent dependency chains per thread; 148*8 CTAs of 256 threads.
// Prints warp-instructions per clock per SM (4 = one per scheduler per clock).
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

constexpr int ITER = 512;

#define CHAINS8(OP)                                                              \
    _Pragma("unroll 1") for (int it = 0; it < ITER; ++it) {                       \
        _Pragma("unroll") for (int r = 0; r < 4; ++r) {                           \
            OP(a0, a1, a2) OP(a1, a2, a3) OP(a2, a3, a4) OP(a3, a4, a5) OP(a4, a5, a6) OP(a5, a6, a7) OP(a6, a7, a0) OP(a7, a0, a1) \
        }                                                                         \
    }

#define KERNEL(name, OPA)                                                                        \
    __global__ void name(unsigned* out, unsigned x, unsigned y, long long* cyc) {                \
        unsigned a0 = threadIdx.x, a1 = a0 * 3 + x, a2 = a0 * 5 + x, a3 = a0 * 7 + x,           \
                 a4 = a0 * 11 + y, a5 = a0 * 13 + y, a6 = a0 * 17 + y, a7 = a0 * 19 + y;         \
        long long t0 = clock64();                                                                \
        CHAINS8(OPA)                                                                             \
        long long t1 = clock64();                                                                \
        out[blockIdx.x * blockDim.x + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;      \
        if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;                                         \
    }

__device__ __forceinline__ unsigned h2max(unsigned a, unsigned b) {
    unsigned r;
    asm volatile("max.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
__device__ __forceinline__ unsigned h2add(unsigned a, unsigned b) {
    unsigned r;
    asm volatile("add.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}

#define OP_VMAX(a, b, c) a = __vmaxu2(a, b);
#define OP_VMIN3(a, b, c) a = __vimin3_u16x2(a, b, c);
#define OP_HMAX(a, b, c) a = h2max(a, b);
#define OP_PRMT(a, b, c) a = __byte_perm(a, b, 0x5140);
#define OP_LOP(a, b, c) a = (a & b) ^ c;
#define OP_IADD3(a, b, c) a = a + b + c;
#define OP_IMAD(a, b, c) a = a * b + c;
#define OP_SHF(a, b, c) a = __funnelshift_r(a, b, 8);
#define OP_ABSD(a, b, c) a = __vabsdiffu4(a, b);
#define OP_POPC(a, b, c) a = __popc(a) + b;
#define OP_FLO(a, b, c) a = __clz(a) + b;
#define OP_MIX_VH(a, b, c) a = __vmaxu2(a, b); a = h2max(a, c);
#define OP_MIX_VI(a, b, c) a = __vmaxu2(a, b); a = a * b + c;
#define OP_MIX_V3H(a, b, c) a = __vimin3_u16x2(a, b, c); a = h2max(a, c);
#define OP_MIX_PH(a, b, c) a = __byte_perm(a, b, 0x5140); a = h2max(a, c);
#define OP_MIX_VP(a, b, c) a = __vmaxu2(a, b); a = __byte_perm(a, c, 0x5140);
#define OP_MIX_HI(a, b, c) a = h2max(a, b); a = a * b + c;
#define OP_MIX_HA(a, b, c) a = h2max(a, b); a = h2add(a, c);

#define OP_R31(a, b, c) a = __vmaxu2(a, b); a = __vminu2(a, c); a = __vmaxu2(a, b); a = a * b + c;
#define OP_R21(a, b, c) a = __vmaxu2(a, b); a = __vminu2(a, c); a = a * b + c;
#define OP_R12(a, b, c) a = __vmaxu2(a, b); a = a * b + c; a = a * c + b;
#define OP_R13(a, b, c) a = __vmaxu2(a, b); a = a * b + c; a = a * c + b; a = a * b + b;
#define OP_LDSV(a, b, c) a = __vmaxu2(a, b); a = sm[(a & 1023)];
#define OP_FADD(a, b, c) a = __float_as_uint(__uint_as_float(a) + __uint_as_float(b));
#define OP_VF(a, b, c) a = __vmaxu2(a, b); a = __float_as_uint(__uint_as_float(a) + __uint_as_float(c));
#define OP_VH(a, b, c) a = __vmaxu2(a, b); a = h2add(a, c);
#define OP_PRMTI(a, b, c) a = __byte_perm(a, b, 0x5140); a = a * b + c;
#define OP_LOPI(a, b, c) a = (a & b) ^ c; a = a * b + c;
#define OP_IMADHI(a, b, c) a = __umulhi(a, b) ^ c;
#define OP_IDP4(a, b, c) a = __dp4a(a, b, c);
#define OP_IDP2(a, b, c) a = __dp2a_lo(a, b, c);
#define OP_F2I(a, b, c) a = (unsigned)__float2int_rn(__uint_as_float(a)) + b;
#define OP_I2F(a, b, c) a = __float_as_uint((float)(int)a) ^ b;
#define OP_FFMA(a, b, c) a = __float_as_uint(fmaf(__uint_as_float(a), __uint_as_float(b), __uint_as_float(c)));
KERNEL(k_imadhi, OP_IMADHI)
KERNEL(k_idp4, OP_IDP4)
KERNEL(k_idp2, OP_IDP2)
KERNEL(k_f2i, OP_F2I)
KERNEL(k_i2f, OP_I2F)
KERNEL(k_ffma, OP_FFMA)
KERNEL(k_r31, OP_R31)
KERNEL(k_r21, OP_R21)
KERNEL(k_r12, OP_R12)
KERNEL(k_r13, OP_R13)
KERNEL(k_fadd, OP_FADD)
KERNEL(k_vf, OP_VF)
KERNEL(k_vh, OP_VH)
KERNEL(k_prmti, OP_PRMTI)
KERNEL(k_lopi, OP_LOPI)
KERNEL(k_vmax, OP_VMAX)
KERNEL(k_vmin3, OP_VMIN3)
KERNEL(k_hmax, OP_HMAX)
KERNEL(k_prmt, OP_PRMT)
KERNEL(k_lop, OP_LOP)
KERNEL(k_iadd3, OP_IADD3)
KERNEL(k_imad, OP_IMAD)
KERNEL(k_shf, OP_SHF)
KERNEL(k_absd, OP_ABSD)
KERNEL(k_popc, OP_POPC)
KERNEL(k_flo, OP_FLO)
KERNEL(k_mix_vh, OP_MIX_VH)
KERNEL(k_mix_vi, OP_MIX_VI)
KERNEL(k_mix_v3h, OP_MIX_V3H)
KERNEL(k_mix_ph, OP_MIX_PH)
KERNEL(k_mix_vp, OP_MIX_VP)
KERNEL(k_mix_hi, OP_MIX_HI)
KERNEL(k_mix_ha, OP_MIX_HA)

// shared-memory and shuffle issue rates
__global__ void k_lds(unsigned* out, unsigned x, unsigned y, long long* cyc) {
    __shared__ unsigned s[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) s[i] = i * x;
    __syncthreads();
    unsigned a0 = threadIdx.x & 31, a1 = a0 + 32, a2 = a0 + 64, a3 = a0 + 96, acc = 0;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            a0 = s[a0 & 2047]; a1 = s[a1 & 2047]; a2 = s[a2 & 2047]; a3 = s[a3 & 2047];
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3 ^ acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_shfl(unsigned* out, unsigned x, unsigned y, long long* cyc) {
    unsigned a0 = threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, a4 = a0 + y, a5 = a0 ^ y, a6 = a0 * y, a7 = a0 - y;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            a0 = __shfl_xor_sync(0xffffffffu, a0, 1); a1 = __shfl_xor_sync(0xffffffffu, a1, 2);
            a2 = __shfl_xor_sync(0xffffffffu, a2, 4); a3 = __shfl_xor_sync(0xffffffffu, a3, 8);
            a4 = __shfl_xor_sync(0xffffffffu, a4, 1); a5 = __shfl_xor_sync(0xffffffffu, a5, 2);
            a6 = __shfl_xor_sync(0xffffffffu, a6, 4); a7 = __shfl_xor_sync(0xffffffffu, a7, 8);
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

typedef void (*kern_t)(unsigned*, unsigned, unsigned, long long*);

int main() {
    const int blocks = 148 * 8, threads = 256;
    unsigned* out;
    long long* cyc;
    cudaMalloc(&out, blocks * threads * 4);
    cudaMalloc(&cyc, blocks * 8);
    long long* h = (long long*)malloc(blocks * 8);
    struct { const char* name; kern_t k; int perIter; } tests[] = {
        {"VIMNMX.U16x2", k_vmax, 32},   {"VIMNMX3.U16x2", k_vmin3, 32}, {"HMNMX2", k_hmax, 32},
        {"PRMT", k_prmt, 32},           {"LOP3", k_lop, 32},            {"IADD3", k_iadd3, 32},
        {"IMAD", k_imad, 32},           {"SHF (funnel)", k_shf, 32},    {"VABSDIFF4", k_absd, 32},
        {"POPC+IADD", k_popc, 64},      {"FLO+IADD", k_flo, 64},        {"VIMNMX + HMNMX2", k_mix_vh, 64},
        {"VIMNMX + IMAD", k_mix_vi, 64}, {"VIMNMX3 + HMNMX2", k_mix_v3h, 64}, {"PRMT + HMNMX2", k_mix_ph, 64},
        {"VIMNMX + PRMT", k_mix_vp, 64}, {"HMNMX2 + IMAD", k_mix_hi, 64}, {"HMNMX2 + HADD2", k_mix_ha, 64},
        {"LDS.32 (conflict-free)", k_lds, 32}, {"SHFL", k_shfl, 32},
        {"3 VIMNMX : 1 IMAD", k_r31, 128}, {"2 VIMNMX : 1 IMAD", k_r21, 96}, {"1 VIMNMX : 2 IMAD", k_r12, 96}, {"1 VIMNMX : 3 IMAD", k_r13, 128},
        {"IMAD.HI + LOP3", k_imadhi, 64}, {"IDP.4A", k_idp4, 32}, {"IDP.2A", k_idp2, 32}, {"F2I + IADD", k_f2i, 64},
        {"I2F + LOP3", k_i2f, 64}, {"FFMA", k_ffma, 32},
        {"FADD", k_fadd, 32}, {"VIMNMX + FADD", k_vf, 64}, {"VIMNMX + HADD2", k_vh, 64}, {"PRMT + IMAD", k_prmti, 64}, {"LOP3 + IMAD", k_lopi, 64},
    };
    for (auto& t : tests) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        t.k<<<blocks, threads>>>(out, 0x01000100u, 0x00030005u, cyc);
        cudaEventRecord(e0);
        t.k<<<blocks, threads>>>(out, 0x01000100u, 0x00030005u, cyc);
        cudaEventRecord(e1);
        cudaDeviceSynchronize();
        float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
        cudaMemcpy(h, cyc, blocks * 8, cudaMemcpyDeviceToHost);
        double avg = 0;
        for (int i = 0; i < blocks; ++i) avg += (double)h[i];
        avg /= blocks;
        // 8 CTAs of 8 warps per SM resident = 64 warps; each executes ITER * perIter instructions in `avg` cycles
        const double perSm = 64.0 * ITER * t.perIter / avg;
        // wall-clock rate: all 148 SMs, 64 warps each
        const double perSmWall = 64.0 * ITER * t.perIter / (ms * 1e-3 * 1.965e9);
        printf("%-26s %6.2f warp-instr/clk/SM by clock64 (%.0f cycles per CTA), %6.2f by wall clock at 1.965 GHz (%.3f ms)\n", t.name, perSm, avg, perSmWall, ms);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
