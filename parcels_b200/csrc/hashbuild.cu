// hashbuild.cu -- device-side construction of the spatial-hash table of curvilinear grids (reference
// _core/spatialhash.py:269-387, `_initialize_hash_table`): one (Morton key, face) entry per hash cell of every face's
// quantised bounding box, sorted by (key, face), stored CSR (unique keys ascending, start / count per key, flat face ids
// ascending within a key) -- so that the device query walks candidates in the reference's order.
//
// Input: per face the quantised bounding box packed 6 x 10 bits (xlo | xhi<<10 | ylo<<20 | yhi<<30 | zlo<<40 | zhi<<50;
// an invalid face has xlo > xhi).  The boxes come from the float part of the reference's constructor (:45-228: unit-sphere
// xyz in the coordinate dtype, per-face min/max, quantisation, bitwidth budget search), which stays on the host in NumPy:
// it is O(faces), and libm's sin/cos are what make the reference's boxes -- the table is then a pure integer function of
// them, bit-identical to the reference's.  The O(entries) part (expansion, Morton encode, 64-bit sort, CSR; 12.7 M entries
// for ORCA025, 11 s of NumPy) runs here: count -> exclusive scan -> expand -> radix sort -> head flags -> scan -> CSR.
// cub's DeviceScan / DeviceRadixSort do the two primitive passes (one-time grid setup, not the hot path).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "common.cuh"

__device__ __forceinline__ unsigned hb_dilate(unsigned n) {  // 10 bits -> every third bit (_encode_quantized_morton3d)
    n &= 0x3FFu;
    n = (n | (n << 16)) & 0xFF0000FFu;
    n = (n | (n << 8)) & 0x0300F00Fu;
    n = (n | (n << 4)) & 0x030C30C3u;
    n = (n | (n << 2)) & 0x09249249u;
    return n;
}

struct QBox {
    int xlo, xhi, ylo, yhi, zlo, zhi;
};
__device__ __forceinline__ QBox hb_unpack(unsigned long long q) {
    return QBox{(int)(q & 1023), (int)((q >> 10) & 1023), (int)((q >> 20) & 1023), (int)((q >> 30) & 1023), (int)((q >> 40) & 1023),
                (int)((q >> 50) & 1023)};
}

__global__ void hb_count(const unsigned long long* __restrict__ qbox, long long nfaces, long long* __restrict__ counts) {
    const long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nfaces) return;
    const QBox b = hb_unpack(qbox[f]);
    const bool valid = b.xlo <= b.xhi && b.ylo <= b.yhi && b.zlo <= b.zhi;
    counts[f] = valid ? (long long)(b.xhi - b.xlo + 1) * (b.yhi - b.ylo + 1) * (b.zhi - b.zlo + 1) : 0;
}

// entry e of the face-major enumeration (:318-336): face = last f with offsets[f] <= e; intra = e - offsets[face];
// cell = (xlo + intra / (ny nz), ylo + (intra % (ny nz)) / nz, zlo + (intra % (ny nz)) % nz)
__global__ void hb_expand(const unsigned long long* __restrict__ qbox, const long long* __restrict__ offsets, long long nfaces,
                          long long nent, unsigned long long* __restrict__ packed) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nent) return;
    long long lo = 0, hi = nfaces;  // first f with offsets[f] > e
    while (lo < hi) {
        const long long m = (lo + hi) >> 1;
        if (offsets[m] <= e) lo = m + 1; else hi = m;
    }
    const long long f = lo - 1;
    const QBox b = hb_unpack(qbox[f]);
    const long long intra = e - offsets[f];
    const int ny = b.yhi - b.ylo + 1, nz = b.zhi - b.zlo + 1;
    const long long nynz = (long long)ny * nz;
    const int cx = b.xlo + (int)(intra / nynz);
    const int rem = (int)(intra % nynz);
    const int cy = b.ylo + rem / nz, cz = b.zlo + rem % nz;
    const unsigned code = (hb_dilate((unsigned)cz) << 2) | (hb_dilate((unsigned)cy) << 1) | hb_dilate((unsigned)cx);
    packed[e] = ((unsigned long long)code << 32) | (unsigned long long)(unsigned)f;
}

__global__ void hb_heads(const unsigned long long* __restrict__ sorted, long long nent, int* __restrict__ head, unsigned* __restrict__ faces) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nent) return;
    const unsigned long long p = sorted[e];
    faces[e] = (unsigned)p;
    head[e] = (e == 0 || (unsigned)(sorted[e - 1] >> 32) != (unsigned)(p >> 32)) ? 1 : 0;
}

__global__ void hb_csr(const unsigned long long* __restrict__ sorted, const int* __restrict__ head, const int* __restrict__ kidx,
                       long long nent, unsigned* __restrict__ keys, long long* __restrict__ starts) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nent || !head[e]) return;
    const int k = kidx[e];
    keys[k] = (unsigned)(sorted[e] >> 32);
    starts[k] = e;
}

__global__ void hb_counts(const long long* __restrict__ starts, long long nkeys, long long nent, long long* __restrict__ counts) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nkeys) return;
    counts[k] = (k + 1 < nkeys ? starts[k + 1] : nent) - starts[k];
}

// bucket[b] = first key index with (key >> shift) >= b, b = 0 .. 2^bits
__global__ void hb_bucket(const unsigned* __restrict__ keys, long long nkeys, int shift, long long nb, int* __restrict__ bucket) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b > nb) return;
    const unsigned long long target = (unsigned long long)b << shift;
    long long lo = 0, hi = nkeys;
    while (lo < hi) {
        const long long m = (lo + hi) >> 1;
        if ((unsigned long long)keys[m] < target) lo = m + 1; else hi = m;
    }
    bucket[b] = (int)lo;
}

#define HB_CK(x)                       \
    do {                               \
        cudaError_t ce_ = (x);         \
        if (ce_ != cudaSuccess) {      \
            hb_free(t, tmp, ntmp);     \
            return ce_;                \
        }                              \
    } while (0)

static void hb_free(HashTableDev& t, void** tmp, int ntmp) {
    for (int k = 0; k < ntmp; ++k)
        if (tmp[k]) cudaFree(tmp[k]);
    void* own[5] = {t.keys, t.starts, t.counts, t.faces, t.bucket};
    for (void* p : own)
        if (p) cudaFree(p);
    t = HashTableDev{};
}

// Builds the table in freshly cudaMalloc'ed buffers (ownership passes to the caller on success).
cudaError_t build_hash_table_device(const unsigned long long* d_qbox, long long nfaces, int bucket_bits, HashTableDev& t, cudaStream_t s) {
    t = HashTableDev{};
    enum { OFFS, CNT, PACK_A, PACK_B, TEMP, HEAD, KIDX, NTMP };
    void* tmp[NTMP] = {};
    const int ntmp = NTMP;
    const int B = 256;
    HB_CK(cudaMalloc(&tmp[CNT], (size_t)nfaces * 8));
    HB_CK(cudaMalloc(&tmp[OFFS], (size_t)nfaces * 8));
    long long* counts = (long long*)tmp[CNT];
    long long* offsets = (long long*)tmp[OFFS];
    hb_count<<<(unsigned)((nfaces + B - 1) / B), B, 0, s>>>(d_qbox, nfaces, counts);
    HB_CK(cudaGetLastError());
    size_t tb = 0;
    HB_CK(cub::DeviceScan::ExclusiveSum(nullptr, tb, counts, offsets, (int)nfaces, s));
    HB_CK(cudaMalloc(&tmp[TEMP], tb ? tb : 1));
    HB_CK(cub::DeviceScan::ExclusiveSum(tmp[TEMP], tb, counts, offsets, (int)nfaces, s));
    long long last[2] = {0, 0};
    HB_CK(cudaMemcpyAsync(&last[0], offsets + nfaces - 1, 8, cudaMemcpyDeviceToHost, s));
    HB_CK(cudaMemcpyAsync(&last[1], counts + nfaces - 1, 8, cudaMemcpyDeviceToHost, s));
    HB_CK(cudaStreamSynchronize(s));
    const long long nent = last[0] + last[1];
    if (nent < 1 || nent > INT_MAX) {
        hb_free(t, tmp, ntmp);
        return cudaErrorInvalidValue;
    }
    cudaFree(tmp[TEMP]); tmp[TEMP] = nullptr;
    HB_CK(cudaMalloc(&tmp[PACK_A], (size_t)nent * 8));
    HB_CK(cudaMalloc(&tmp[PACK_B], (size_t)nent * 8));
    unsigned long long* pa = (unsigned long long*)tmp[PACK_A];
    unsigned long long* pb = (unsigned long long*)tmp[PACK_B];
    hb_expand<<<(unsigned)((nent + B - 1) / B), B, 0, s>>>(d_qbox, offsets, nfaces, nent, pa);
    HB_CK(cudaGetLastError());
    // (code << 32 | face): unsigned order = by code, ties by ascending face id (:346-358); 30 + 32 significant bits
    HB_CK(cub::DeviceRadixSort::SortKeys(nullptr, tb, pa, pb, (int)nent, 0, 62, s));
    HB_CK(cudaMalloc(&tmp[TEMP], tb ? tb : 1));
    HB_CK(cub::DeviceRadixSort::SortKeys(tmp[TEMP], tb, pa, pb, (int)nent, 0, 62, s));
    HB_CK(cudaStreamSynchronize(s));
    cudaFree(tmp[TEMP]); tmp[TEMP] = nullptr;
    cudaFree(tmp[PACK_A]); tmp[PACK_A] = nullptr;
    HB_CK(cudaMalloc((void**)&t.faces, (size_t)nent * 4));
    HB_CK(cudaMalloc(&tmp[HEAD], (size_t)nent * 4));
    HB_CK(cudaMalloc(&tmp[KIDX], (size_t)nent * 4));
    int* head = (int*)tmp[HEAD];
    int* kidx = (int*)tmp[KIDX];
    hb_heads<<<(unsigned)((nent + B - 1) / B), B, 0, s>>>(pb, nent, head, t.faces);
    HB_CK(cudaGetLastError());
    HB_CK(cub::DeviceScan::ExclusiveSum(nullptr, tb, head, kidx, (int)nent, s));
    HB_CK(cudaMalloc(&tmp[TEMP], tb ? tb : 1));
    HB_CK(cub::DeviceScan::ExclusiveSum(tmp[TEMP], tb, head, kidx, (int)nent, s));
    int lastk[2] = {0, 0};
    HB_CK(cudaMemcpyAsync(&lastk[0], kidx + nent - 1, 4, cudaMemcpyDeviceToHost, s));
    HB_CK(cudaMemcpyAsync(&lastk[1], head + nent - 1, 4, cudaMemcpyDeviceToHost, s));
    HB_CK(cudaStreamSynchronize(s));
    const long long nkeys = (long long)lastk[0] + lastk[1];
    HB_CK(cudaMalloc((void**)&t.keys, (size_t)nkeys * 4));
    HB_CK(cudaMalloc((void**)&t.starts, (size_t)nkeys * 8));
    HB_CK(cudaMalloc((void**)&t.counts, (size_t)nkeys * 8));
    hb_csr<<<(unsigned)((nent + B - 1) / B), B, 0, s>>>(pb, head, kidx, nent, t.keys, t.starts);
    HB_CK(cudaGetLastError());
    hb_counts<<<(unsigned)((nkeys + B - 1) / B), B, 0, s>>>(t.starts, nkeys, nent, t.counts);
    HB_CK(cudaGetLastError());
    const long long nb = 1LL << bucket_bits;
    HB_CK(cudaMalloc((void**)&t.bucket, (size_t)(nb + 1) * 4));
    hb_bucket<<<(unsigned)((nb + 1 + B - 1) / B), B, 0, s>>>(t.keys, nkeys, 30 - bucket_bits, nb, t.bucket);
    HB_CK(cudaGetLastError());
    HB_CK(cudaStreamSynchronize(s));
    t.nkeys = nkeys;
    t.nent = nent;
    for (int k = 0; k < ntmp; ++k)
        if (tmp[k]) cudaFree(tmp[k]);
    return cudaSuccess;
}
