// 7x7 sigma-2 Gaussian blur of every pyramid level (north-star kernel 5): ORBextractor.cc:1103-1104,
// cv::GaussianBlur(level, Size(7,7), 2, 2, BORDER_REFLECT_101) on CV_8UC1.
//
// OpenCV's 8-bit path is fixed point and separable: kernel [18 34 48 56 48 34 18]/256 per axis, horizontal sums exact
// (<= 255*256, 16 bits), vertical sums exact in 32 bits, one rounding (v + 2^15) >> 16.  The reference blurs a clone
// of the level, so the border is the level's own reflection -- which is exactly what the pyramid's 19-px frame already
// holds, so the halo is read straight from the padded level: no border logic here.
//
// The kernel is instruction-bound, not DRAM-bound, so it is built around registers and the integer dot-product unit:
//   * a work item is a 4-pixel-wide column (one 32-bit word per row) walking down 32 output rows; items are numbered
//     band-major inside a level and laid over the threads linearly (full warps whatever the level's width);
//   * horizontal pass: per input row three aligned words; a pixel's 8 neighbourhood bytes are two funnel shifts, its sum
//     two IDP.4A (bytes x 8-bit taps);
//   * vertical pass: the sums of two consecutive input rows are packed into one 16x2 register per pixel, so a 7-tap
//     column is four IDP.2A (16-bit sums x 8-bit taps) over the last four row pairs; the accumulator starts at 2^15 so
//     the rounding is free, and the four result bytes are picked with PRMT.
// blur_kernel (small calls): no shared memory, no barriers; all levels in one launch.
// blur_staged_kernel (batches, round 2): the same arithmetic with a band's input rows staged in shared memory by one
// cp.async.bulk, one launch per level -- see the comment at the kernel.
#include <algorithm>
#include <cstdlib>

#include "extractor.h"

namespace orbb {

constexpr int BL_ROWS = 32;                 // output rows per work item (even)
constexpr int BL_THREADS = 128;

// taps as dot-product operands (byte 0 first)
constexpr unsigned int TAP_H_LO = 0x38302212u;   // 18 34 48 56   x  p[x-3] p[x-2] p[x-1] p[x]
constexpr unsigned int TAP_H_HI = 0x00122230u;   // 48 34 18  0   x  p[x+1] p[x+2] p[x+3] p[x+4]
constexpr unsigned int TAP_V_E0 = 0x38302212u;   // even output row: row pairs weigh (18,34) (48,56)
constexpr unsigned int TAP_V_E1 = 0x00122230u;   //                                   (48,34) (18, 0)
constexpr unsigned int TAP_V_O0 = 0x30221200u;   // odd output row:                   ( 0,18) (34,48)
constexpr unsigned int TAP_V_O1 = 0x12223038u;   //                                   (56,48) (34,18)

// Horizontal sums of one input row for four adjacent pixels from the three aligned words that hold bytes x-4 .. x+7:
// pixel k needs bytes k+1 .. k+7 of that window.  The kernel's busiest pipe is the FMA-heavy one (IDP issues there every
// other clock), so the row costs as few dot products as possible: pixels 0 and 3 take the unshifted words with shifted
// TAP vectors (zero outside the pixel's seven bytes), pixels 1 and 2 share ONE window shifted by two bytes (two funnel
// shifts on the ALU pipe) with the two tap alignments -- 8 IDP.4A + 2 SHF per 4 pixels (integer sums: same result as
// any other split; 10 IDP.4A without shifts was 2.70 ms per 4096 frames, 6 shifts + 8 IDP.4A more instructions).
__device__ __forceinline__ void blur_hrow(unsigned int w0, unsigned int w1, unsigned int w2, unsigned int (&h)[4]) {
#ifdef BL_HROW10
    h[0] = __dp4a(w0, 0x30221200u, __dp4a(w1, 0x12223038u, 0u));                              // bytes 1..3 | 4..7
    h[1] = __dp4a(w0, 0x22120000u, __dp4a(w1, 0x22303830u, __dp4a(w2, 0x00000012u, 0u)));     // bytes 2..3 | 4..7 | 8
    h[2] = __dp4a(w0, 0x12000000u, __dp4a(w1, 0x30383022u, __dp4a(w2, 0x00001222u, 0u)));     // byte 3 | 4..7 | 8..9
    h[3] = __dp4a(w1, TAP_H_LO, __dp4a(w2, TAP_H_HI, 0u));                                    // bytes 4..7 | 8..10
#else
    const unsigned int X = __funnelshift_r(w0, w1, 16), Y = __funnelshift_r(w1, w2, 16);      // bytes 2..5, 6..9
    h[0] = __dp4a(w0, 0x30221200u, __dp4a(w1, 0x12223038u, 0u));                              // bytes 1..3 | 4..7
    h[1] = __dp4a(X, TAP_H_LO, __dp4a(Y, TAP_H_HI, 0u));                                      // bytes 2..5 | 6..8
    h[2] = __dp4a(X, 0x30221200u, __dp4a(Y, 0x12223038u, 0u));                                // bytes 3..5 | 6..9
    h[3] = __dp4a(w1, TAP_H_LO, __dp4a(w2, TAP_H_HI, 0u));                                    // bytes 4..7 | 8..10
#endif
}

int blur_cta_count(int w, int h) { return ceil_div(((w + 3) / 4) * ceil_div(h, BL_ROWS), BL_THREADS); }

__global__ void __launch_bounds__(BL_THREADS) blur_kernel(const __grid_constant__ ExtractParams P, const BlurTile* __restrict__ tiles) {
    const BlurTile t = tiles[blockIdx.x];
    const LevelGeom& L = P.lv[t.level];
    const int frame = blockIdx.y;
    const int groups = (L.w + 3) >> 2;
    const int item = t.cta * BL_THREADS + threadIdx.x;
    const int band = item / groups, g = item - band * groups;
    const int x0 = 4 * g, y0 = band * BL_ROWS;
    if (y0 >= L.h) return;
    const int pitch = L.pitch, bpitch = L.bpitch;
    const int lastRow = L.h + 2 * kEdge - 1;
    // input row j of this item is buffer row y0 - 3 + j + kEdge (frame rows included), clamped below the last band
    const unsigned char* col = P.pyr + (size_t)frame * P.pyrFrameBytes + L.pyrOff + kPadLeft + x0 - 4;
    unsigned char* out = P.blur + (size_t)frame * P.blurFrameBytes + L.blurOff + (size_t)y0 * bpitch + x0;
    const int rowsOut = min(BL_ROWS, L.h - y0);

    // the three aligned words of input row j that hold the item's pixels and their 3-px neighbourhood
    auto load_row = [&](int j, unsigned int (&w)[3]) {
        const unsigned int* p = reinterpret_cast<const unsigned int*>(col + (size_t)min(y0 - 3 + j + kEdge, lastRow) * pitch);
        w[0] = __ldg(p); w[1] = __ldg(p + 1); w[2] = __ldg(p + 2);
    };
    // horizontal sums of one input row for the item's four pixels
    auto hrow = [&](const unsigned int (&w)[3], unsigned int (&h)[4]) { blur_hrow(w[0], w[1], w[2], h); };
    // row pair = input rows 2i, 2i+1 packed per pixel (even row in the low half)
    auto make_pair = [&](const unsigned int (&we)[3], const unsigned int (&wo)[3], unsigned int (&pr)[4]) {
        unsigned int he[4], ho[4];
        hrow(we, he);
        hrow(wo, ho);
#pragma unroll
        for (int k = 0; k < 4; ++k) pr[k] = __byte_perm(he[k], ho[k], 0x5410);
    };

    unsigned int p0[4], p1[4], p2[4], p3[4];
    unsigned int we[3], wo[3];
    {
        unsigned int a0[3], a1[3], a2[3], a3[3], a4[3], a5[3];
        load_row(0, a0); load_row(1, a1); load_row(2, a2); load_row(3, a3); load_row(4, a4); load_row(5, a5);
        load_row(6, we); load_row(7, wo);
        make_pair(a0, a1, p0);
        make_pair(a2, a3, p1);
        make_pair(a4, a5, p2);
    }
    // pair i completes output rows 2i-6 (input rows 2i-6 .. 2i) and 2i-5 (input rows 2i-5 .. 2i+1); the words of pair i+1
    // are requested before pair i is consumed
#pragma unroll 4
    for (int i = 3; i < BL_ROWS / 2 + 3; ++i) {
        const int o = 2 * i - 6;
        if (o >= rowsOut) break;
        make_pair(we, wo, p3);
        load_row(2 * i + 2, we);
        load_row(2 * i + 3, wo);
        unsigned int ve[4], vo[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            ve[k] = __dp2a_lo(p0[k], TAP_V_E0, __dp2a_hi(p1[k], TAP_V_E0, __dp2a_lo(p2[k], TAP_V_E1, __dp2a_hi(p3[k], TAP_V_E1, 32768u))));
            vo[k] = __dp2a_lo(p0[k], TAP_V_O0, __dp2a_hi(p1[k], TAP_V_O0, __dp2a_lo(p2[k], TAP_V_O1, __dp2a_hi(p3[k], TAP_V_O1, 32768u))));
            p0[k] = p1[k]; p1[k] = p2[k]; p2[k] = p3[k];
        }
        // byte 2 of each sum is (v >> 16) & 0xff (sums stay below 2^24)
        unsigned int* dst = reinterpret_cast<unsigned int*>(out + (size_t)o * bpitch);
        dst[0] = __byte_perm(__byte_perm(ve[0], ve[1], 0x0062), __byte_perm(ve[2], ve[3], 0x0062), 0x5410);
        if (o + 1 < rowsOut)
            *reinterpret_cast<unsigned int*>(out + (size_t)(o + 1) * bpitch) =
                __byte_perm(__byte_perm(vo[0], vo[1], 0x0062), __byte_perm(vo[2], vo[3], 0x0062), 0x5410);
    }
}

// ---- batches: the same arithmetic with the input rows staged through shared memory --------------------------------
// blur_kernel requests every input byte three times (a thread's three words overlap its neighbours') and pays 64-bit
// address arithmetic and a row clamp per input row.  Here a CTA owns a band of 32 output rows over the full width of one
// level of one frame: the 38 input rows of the band are CONTIGUOUS in the padded pyramid level, one cp.async.bulk brings
// them into shared memory (each byte once, no registers held), and thread g walks down its 4-pixel column with 32-bit
// shared-memory addresses.  One buffer per CTA and several CTAs per SM: while one waits for its copy the others compute.
// Persistent CTAs stride over (frame, band); one launch per level (block size = the level's column groups).
__global__ void __launch_bounds__(512)
blur_staged_kernel(const __grid_constant__ ExtractParams P, int level, int nBands, int nTiles, int nBuf, int bufBytes) {
    extern __shared__ __align__(128) unsigned char bsm[];
    const LevelGeom& L = P.lv[level];
    const int tid = threadIdx.x;
    const int groups = (L.w + 3) >> 2;
    const int pitch = L.pitch, bpitch = L.bpitch;
    const unsigned int bar0 = (unsigned int)__cvta_generic_to_shared(bsm), tile0 = bar0 + 128;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const bool active = tid < groups;
    const int issuerTid = (int)blockDim.x > groups ? (int)blockDim.x - 1 : 0;     // a thread without pixels, if the block has one
    auto issue = [&](int t, int buf) {      // the band's input rows y0-3 .. y0+rowsOut+2 = padded rows y0+16 .. ; rowsOut + 6 of them
        if (tid == issuerTid) {
            const int frame = t / nBands, band = t - frame * nBands;
            const int y0 = band * BL_ROWS;
            const int rowsOut = min(BL_ROWS, L.h - y0);
            const unsigned char* src = P.pyr + (size_t)frame * P.pyrFrameBytes + L.pyrOff + (size_t)(y0 + kEdge - 3) * pitch;
            const unsigned int bytes = (unsigned int)(rowsOut + 6) * (unsigned int)pitch;
            const unsigned int bar = bar0 + 8 * buf;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             tile0 + buf * bufBytes),
                         "l"(src), "r"(bytes), "r"(bar)
                         : "memory");
        }
    };
    // nBuf == 2: the next band's rows are in flight while this band is computed; nBuf == 1: one buffer, twice the CTAs per SM
    if (nBuf == 2 && (int)blockIdx.x < nTiles) issue(blockIdx.x, 0);
    int it = 0;
    for (int t = blockIdx.x; t < nTiles; t += gridDim.x, ++it) {
        const int frame = t / nBands, band = t - frame * nBands;
        const int y0 = band * BL_ROWS;
        const int rowsOut = min(BL_ROWS, L.h - y0);
        const int buf = nBuf == 2 ? (it & 1) : 0;
        if (nBuf == 2) {
            if (t + (int)gridDim.x < nTiles) issue(t + gridDim.x, buf ^ 1);
        } else {
            issue(t, 0);
        }
        const unsigned int bar = bar0 + 8 * buf, parity = (unsigned int)(nBuf == 2 ? it >> 1 : it) & 1u;
        const unsigned int colAddr = tile0 + buf * bufBytes + kPadLeft + 4 * tid - 4;   // word of pixels 4g-4 .. 4g-1 of the first input row
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "BL_WAIT_%=:\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
            "@p bra BL_DONE_%=;\n"
            "bra BL_WAIT_%=;\n"
            "BL_DONE_%=:\n"
            "}\n" ::"r"(bar),
            "r"(parity)
            : "memory");
        if (active) {
            unsigned char* out = P.blur + (size_t)frame * P.blurFrameBytes + L.blurOff + (size_t)y0 * bpitch + 4 * tid;
            unsigned int rowAddr = colAddr;
            auto load_row = [&](unsigned int (&w)[3]) {
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w[0]) : "r"(rowAddr));
                asm volatile("ld.shared.u32 %0, [%1+4];" : "=r"(w[1]) : "r"(rowAddr));
                asm volatile("ld.shared.u32 %0, [%1+8];" : "=r"(w[2]) : "r"(rowAddr));
                rowAddr += pitch;
            };
            auto make_pair = [&](unsigned int (&pr)[4]) {      // the next two input rows, packed per pixel (even row in the low half)
                unsigned int we[3], wo[3], he[4], ho[4];
                load_row(we);
                load_row(wo);
                blur_hrow(we[0], we[1], we[2], he);
                blur_hrow(wo[0], wo[1], wo[2], ho);
#pragma unroll
                for (int k = 0; k < 4; ++k) pr[k] = __byte_perm(he[k], ho[k], 0x5410);
            };
            unsigned int p0[4], p1[4], p2[4], p3[4];
            make_pair(p0);
            make_pair(p1);
            make_pair(p2);
            // pair i completes output rows 2i-6 and 2i-5; the band's buffer holds rowsOut + 6 input rows, so the odd row of
            // the last pair of an odd-height band is read from the row behind them (inside the buffer's slack) and unused
#pragma unroll 4
            for (int o = 0; o < rowsOut; o += 2) {      // unrolled by the period of the four-pair ring: no register moves
                make_pair(p3);
                unsigned int ve[4], vo[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    ve[k] = __dp2a_lo(p0[k], TAP_V_E0, __dp2a_hi(p1[k], TAP_V_E0, __dp2a_lo(p2[k], TAP_V_E1, __dp2a_hi(p3[k], TAP_V_E1, 32768u))));
                    vo[k] = __dp2a_lo(p0[k], TAP_V_O0, __dp2a_hi(p1[k], TAP_V_O0, __dp2a_lo(p2[k], TAP_V_O1, __dp2a_hi(p3[k], TAP_V_O1, 32768u))));
                    p0[k] = p1[k]; p1[k] = p2[k]; p2[k] = p3[k];
                }
                *reinterpret_cast<unsigned int*>(out) = __byte_perm(__byte_perm(ve[0], ve[1], 0x0062), __byte_perm(ve[2], ve[3], 0x0062), 0x5410);
                if (o + 1 < rowsOut)
                    *reinterpret_cast<unsigned int*>(out + bpitch) =
                        __byte_perm(__byte_perm(vo[0], vo[1], 0x0062), __byte_perm(vo[2], vo[3], 0x0062), 0x5410);
                out += 2 * bpitch;
            }
        }
        __syncthreads();   // every thread is done with the buffer before the next band's copy is issued
    }
}

// shared memory of the staged kernel for a level: barrier block + nBuf buffers of (BL_ROWS + 6 input rows + 1 row of slack)
static int blur_staged_buffers() {
    static const int n = getenv("ORBB_BLUR_NBUF") ? std::max(1, std::min(2, atoi(getenv("ORBB_BLUR_NBUF")))) : 1;   // tuning aid (2 measured slower)
    return n;
}
static int blur_staged_buf_bytes(const LevelGeom& L) { return ((BL_ROWS + 7) * L.pitch + 127) / 128 * 128; }
static int blur_staged_smem(const LevelGeom& L) { return 128 + blur_staged_buffers() * blur_staged_buf_bytes(L); }

int blur_staged_ctas(const LevelGeom& L) {
    const int threads = (((L.w + 3) >> 2) + 31) / 32 * 32, smem = blur_staged_smem(L);
    if (threads > 512 || smem > 200 * 1024) return 0;
    int dev = 0, nSm = 0, perSm = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaFuncSetAttribute(blur_staged_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&nSm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, blur_staged_kernel, threads, smem) != cudaSuccess) return 0;
    return nSm * perSm;
}

int launch_blur(const ExtractParams& P, const BlurTile* dTiles, int nTiles, cudaStream_t st, int* launches) {
    if (nTiles == 0) return ORB_OK;
    if (P.nFrames >= P.pyBulkMinFrames) {
        bool all = true;
        for (int l = 0; l < P.nLevels; ++l) all = all && P.lv[l].blCtas > 0;
        if (all) {
            for (int l = 0; l < P.nLevels; ++l) {
                const LevelGeom& L = P.lv[l];
                const int threads = (((L.w + 3) >> 2) + 31) / 32 * 32, nBands = ceil_div(L.h, BL_ROWS);
                const long long tiles = (long long)nBands * P.nFrames;
                const int grid = (int)std::min<long long>(tiles, (long long)L.blCtas);
                blur_staged_kernel<<<grid, threads, blur_staged_smem(L), st>>>(P, l, nBands, (int)tiles, blur_staged_buffers(),
                                                                               blur_staged_buf_bytes(L));
                ++*launches;
            }
            ORB_CUDA(cudaGetLastError());
            return ORB_OK;
        }
    }
    dim3 grid(nTiles, P.nFrames);
    blur_kernel<<<grid, BL_THREADS, 0, st>>>(P, dTiles);
    ++*launches;
    ORB_CUDA(cudaGetLastError());
    return ORB_OK;
}

}  // namespace orbb
