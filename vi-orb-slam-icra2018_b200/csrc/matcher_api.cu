// C-ABI entry points of the matcher (include/orbb200.h): argument checks, host<->device staging, launches.
#include <cstring>

#include "matcher.h"

using namespace orbb;

extern "C" {

int orbm_create(int device, orbm_handle* out) {
    if (!out) return fail(ORB_ERR_INVALID, "orbm_create: null out");
    *out = nullptr;
    int n = 0;
    ORB_CUDA(cudaGetDeviceCount(&n));
    if (device < 0 || device >= n) return fail(ORB_ERR_INVALID, "orbm_create: device %d of %d", device, n);
    DeviceGuard g(device);
    if (!g.ok) return fail(ORB_ERR_CUDA, "orbm_create: cannot select device %d", device);
    orbm_matcher* m = new orbm_matcher;
    m->device = device;
    cudaError_t e = cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        delete m;
        return fail(ORB_ERR_CUDA, "orbm_create: %s", cudaGetErrorString(e));
    }
    *out = m;
    return ORB_OK;
}

int orbm_destroy(orbm_handle h) {
    if (!h) return ORB_OK;
    DeviceGuard g(h->device);
    cudaStreamSynchronize(h->stream);
    DevBuf* bufs[] = {&h->in0, &h->in1, &h->in2, &h->in3, &h->in4, &h->in5, &h->out0, &h->out1, &h->out2, &h->out3,
                      &h->out4, &h->ws0, &h->ws1, &h->ws2, &h->ws3};
    for (DevBuf* b : bufs) b->release();
    PinnedBuf* pins[] = {&h->pin0, &h->pin1, &h->pin2, &h->pin3, &h->pin4, &h->stage};
    for (PinnedBuf* b : pins) b->release();
    cudaStreamDestroy(h->stream);
    delete h;
    return ORB_OK;
}

int orbm_synchronize(orbm_handle h) {
    if (!h) return fail(ORB_ERR_INVALID, "orbm_synchronize: null handle");
    DeviceGuard g(h->device);
    ORB_CUDA(cudaStreamSynchronize(h->stream));
    return ORB_OK;
}

int orbm_last_launch_count(orbm_handle h, int* n) {
    if (!h || !n) return fail(ORB_ERR_INVALID, "orbm_last_launch_count: null argument");
    *n = h->launches;
    return ORB_OK;
}

int orbm_distance(orbm_handle h, const uint8_t* a, const uint8_t* b, int n, int* dist) {
    ORBM_ENTER(h);
    if (n < 0 || (n > 0 && (!a || !b || !dist))) return fail(ORB_ERR_INVALID, "orbm_distance: bad arguments");
    if (n == 0) return ORB_OK;
    const size_t bytes = (size_t)n * 32;
    ORB_CHECK(h->in0.reserve(bytes));
    ORB_CHECK(h->in1.reserve(bytes));
    ORB_CHECK(h->out0.reserve((size_t)n * 4));
    ORB_CUDA(cudaMemcpyAsync(h->in0.p, a, bytes, cudaMemcpyHostToDevice, h->stream));
    ORB_CUDA(cudaMemcpyAsync(h->in1.p, b, bytes, cudaMemcpyHostToDevice, h->stream));
    ORB_CHECK(launch_distance(h->in0.as<uint8_t>(), h->in1.as<uint8_t>(), n, h->out0.as<int>(), h->stream, &h->launches));
    ORB_CUDA(cudaMemcpyAsync(dist, h->out0.p, (size_t)n * 4, cudaMemcpyDeviceToHost, h->stream));
    ORB_CUDA(cudaStreamSynchronize(h->stream));
    return ORB_OK;
}

int orbm_bruteforce_device(orbm_handle h, const uint8_t* dq, const float* dqa, int nq, const uint8_t* dt,
                           const float* dta, int nt, int nPairs, float ratio, int checkOri, int* dBest, int* dSecond,
                           int* dIdx, int* dM12, int* dN, void* stream) {
    ORBM_ENTER(h);
    if (nq < 0 || nt < 0 || nPairs < 0) return fail(ORB_ERR_INVALID, "orbm_bruteforce: negative size");
    if (nq == 0 || nPairs == 0) return ORB_OK;
    if (!dq || (nt > 0 && !dt) || !dM12 || !dN || (checkOri && (!dqa || (nt > 0 && !dta))))
        return fail(ORB_ERR_INVALID, "orbm_bruteforce: null pointer");
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    const size_t keys = (size_t)nPairs * nq * 4;
    ORB_CHECK(h->ws0.reserve(keys));
    ORB_CHECK(h->ws1.reserve(keys));
    return launch_bruteforce(dq, dqa, nq, dt, dta, nt, nPairs, ratio, checkOri, h->ws0.as<int>(), h->ws1.as<int>(),
                             dBest, dSecond, dIdx, dM12, dN, st, &h->launches);
}

int orbm_bruteforce(orbm_handle h, const uint8_t* q, const float* qa, int nq, const uint8_t* t, const float* ta, int nt,
                    int nPairs, float ratio, int checkOri, int* best, int* second, int* idx, int* m12, int* nmatches) {
    ORBM_ENTER(h);
    if (nq < 0 || nt < 0 || nPairs < 0) return fail(ORB_ERR_INVALID, "orbm_bruteforce: negative size");
    if (nPairs > 0 && !nmatches) return fail(ORB_ERR_INVALID, "orbm_bruteforce: null nmatches");
    if (nq == 0 || nPairs == 0) {
        for (int p = 0; p < nPairs; ++p) nmatches[p] = 0;
        return ORB_OK;
    }
    if (!q || (nt > 0 && !t) || !m12 || !nmatches || (checkOri && (!qa || (nt > 0 && !ta))))
        return fail(ORB_ERR_INVALID, "orbm_bruteforce: null pointer");
    const size_t qb = (size_t)nPairs * nq * 32, tb = (size_t)nPairs * nt * 32;
    const size_t qab = (size_t)nPairs * nq * 4, tab = (size_t)nPairs * nt * 4, ob = (size_t)nPairs * nq * 4;
    ORB_CHECK(h->in0.reserve(qb));
    ORB_CHECK(h->in1.reserve(tb + 32));
    ORB_CHECK(h->in2.reserve(qab));
    ORB_CHECK(h->in3.reserve(tab + 4));
    ORB_CHECK(h->out0.reserve(ob));
    ORB_CHECK(h->out1.reserve(ob));
    ORB_CHECK(h->out2.reserve(ob));
    ORB_CHECK(h->out3.reserve(ob));
    ORB_CHECK(h->out4.reserve((size_t)nPairs * 4));
    cudaStream_t st = h->stream;
    ORB_CUDA(cudaMemcpyAsync(h->in0.p, q, qb, cudaMemcpyHostToDevice, st));
    if (nt > 0) ORB_CUDA(cudaMemcpyAsync(h->in1.p, t, tb, cudaMemcpyHostToDevice, st));
    if (checkOri) {
        ORB_CUDA(cudaMemcpyAsync(h->in2.p, qa, qab, cudaMemcpyHostToDevice, st));
        if (nt > 0) ORB_CUDA(cudaMemcpyAsync(h->in3.p, ta, tab, cudaMemcpyHostToDevice, st));
    }
    int launches = 0;
    const size_t keys = (size_t)nPairs * nq * 4;
    ORB_CHECK(h->ws0.reserve(keys));
    ORB_CHECK(h->ws1.reserve(keys));
    ORB_CHECK(launch_bruteforce(h->in0.as<uint8_t>(), h->in2.as<float>(), nq, h->in1.as<uint8_t>(), h->in3.as<float>(),
                                nt, nPairs, ratio, checkOri, h->ws0.as<int>(), h->ws1.as<int>(), h->out0.as<int>(),
                                h->out1.as<int>(), h->out2.as<int>(), h->out3.as<int>(), h->out4.as<int>(), st,
                                &launches));
    h->launches = launches;
    if (best) ORB_CUDA(cudaMemcpyAsync(best, h->out0.p, ob, cudaMemcpyDeviceToHost, st));
    if (second) ORB_CUDA(cudaMemcpyAsync(second, h->out1.p, ob, cudaMemcpyDeviceToHost, st));
    if (idx) ORB_CUDA(cudaMemcpyAsync(idx, h->out2.p, ob, cudaMemcpyDeviceToHost, st));
    ORB_CUDA(cudaMemcpyAsync(m12, h->out3.p, ob, cudaMemcpyDeviceToHost, st));
    ORB_CUDA(cudaMemcpyAsync(nmatches, h->out4.p, (size_t)nPairs * 4, cudaMemcpyDeviceToHost, st));
    ORB_CUDA(cudaStreamSynchronize(st));
    return ORB_OK;
}

int orbm_allpairs_device(orbm_handle h, const uint8_t* dTable, const float* dAngles, int nKf, int nDesc, int qBegin,
                         int qEnd, int dbBegin, int dbEnd, float ratio, int checkOri, int* dCounts, void* stream) {
    ORBM_ENTER(h);
    if (!dTable || !dCounts || (checkOri && !dAngles)) return fail(ORB_ERR_INVALID, "orbm_allpairs: null pointer");
    cudaStream_t st = stream ? (cudaStream_t)stream : h->stream;
    return launch_allpairs(dTable, dAngles, nKf, nDesc, qBegin, qEnd, dbBegin, dbEnd, ratio, checkOri, dCounts, st,
                           &h->launches);
}

int orbm_distinctive_descriptors(orbm_handle h, const uint8_t* desc, const int* start, int nPoints, int* best,
                                 int* bestMedian) {
    ORBM_ENTER(h);
    if (nPoints < 0 || (nPoints > 0 && (!start || !best))) return fail(ORB_ERR_INVALID, "orbm_distinctive_descriptors: bad arguments");
    if (nPoints == 0) return ORB_OK;
    if (start[0] != 0) return fail(ORB_ERR_INVALID, "orbm_distinctive_descriptors: start[0] must be 0");
    for (int p = 0; p < nPoints; ++p) {
        const int n = start[p + 1] - start[p];
        if (n < 0 || n > 65535)
            return fail(ORB_ERR_INVALID, "orbm_distinctive_descriptors: map point %d has %d observations (0..65535)", p, n);
    }
    const int total = start[nPoints];
    if (total > 0 && !desc) return fail(ORB_ERR_INVALID, "orbm_distinctive_descriptors: null descriptors");
    cudaStream_t st = h->stream;
    ORB_CHECK(h->in0.reserve((size_t)total * 32 + 32));
    ORB_CHECK(h->in1.reserve((size_t)(nPoints + 1) * 4));
    ORB_CHECK(h->out0.reserve((size_t)nPoints * 4));
    ORB_CHECK(h->out1.reserve((size_t)nPoints * 4));
    if (total > 0) ORB_CUDA(cudaMemcpyAsync(h->in0.p, desc, (size_t)total * 32, cudaMemcpyHostToDevice, st));
    ORB_CUDA(cudaMemcpyAsync(h->in1.p, start, (size_t)(nPoints + 1) * 4, cudaMemcpyHostToDevice, st));
    ORB_CHECK(launch_distinctive(h->in0.as<uint8_t>(), h->in1.as<int>(), nPoints, h->out0.as<int>(), h->out1.as<int>(), st,
                                 &h->launches));
    ORB_CUDA(cudaMemcpyAsync(best, h->out0.p, (size_t)nPoints * 4, cudaMemcpyDeviceToHost, st));
    if (bestMedian) ORB_CUDA(cudaMemcpyAsync(bestMedian, h->out1.p, (size_t)nPoints * 4, cudaMemcpyDeviceToHost, st));
    ORB_CUDA(cudaStreamSynchronize(st));
    return ORB_OK;
}

int orbm_popc_peak(orbm_handle h, double* popcPerS) {
    ORBM_ENTER(h);
    if (!popcPerS) return fail(ORB_ERR_INVALID, "orbm_popc_peak: null out");
    h->launches = 5;
    return measure_popc_peak(h->stream, popcPerS);
}

}  // extern "C"
