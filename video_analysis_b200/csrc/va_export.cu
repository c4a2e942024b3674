// va_export.cu -- label images leave the device sparsely.
//
// A label image is 4 bytes per pixel and almost all of it is background: at 1080p with a dozen moving objects about 6 %
// of the 64-pixel chunks hold a foreground pixel.  End to end the chain is bound by PCIe (6.2 MB per frame up, 8.3 MB
// down), so instead of a dense device -> host copy the device compacts the non-empty chunks and writes them, with their
// positions, straight into page-locked host memory (zero-copy stores over PCIe: the host never has to know the size of
// the transfer in advance); the host then rebuilds the dense int32 image the caller gets -- bit-identical to the dense
// copy -- by clearing the chunks the previous user of that buffer left dirty and dropping the new ones in
// (va_host_densify_chunks, threads on the host side of the C ABI).
//
//   chunk        VA_CHUNK_E = 64 consecutive labels of one row (256 bytes); a row of pitch_e elements has
//                ceil(pitch_e / 64) chunks, the last one possibly short
//   chunk id     y * chunks_per_row + c
//   order        raster order per frame (rows, then chunks), so the export is deterministic
//   run chunk    a chunk whose foreground pixels are ONE horizontal run carries one label (a run lies in one component),
//                so it travels as 16 bytes -- (id, label, 64-bit pixel mask) -- instead of 260: inside blobs that is
//                most chunks (1080p chain: 36.7 -> 6.7 MB per 64 frames).  Chunks with several runs travel raw.
#include <cstring>
#include <thread>
#include <vector>

#include "va_device.cuh"

#define VA_CHUNK_E 64
#define EXP_WARPS 8
#define EXP_THREADS (32 * EXP_WARPS)

// the two mask words of chunk c of a row (bits beyond the row's width cleared)
__device__ __forceinline__ unsigned long long exp_chunk_mask(const uint32_t *mrow, int c, int cpr, int words, unsigned lastmask) {
    const int w0 = 2 * c, w1 = 2 * c + 1;
    unsigned lo = (c < cpr && w0 < words) ? mrow[w0] : 0u, hi = (c < cpr && w1 < words) ? mrow[w1] : 0u;
    if (w0 == words - 1) lo &= lastmask;
    if (w1 == words - 1) hi &= lastmask;
    return ((unsigned long long)hi << 32) | lo;
}
// are the set bits of m one run?  (fill the zeros below the lowest set bit: a single run then is 2^k - 1)
__device__ __forceinline__ bool exp_single_run(unsigned long long m) {
    const unsigned long long t = m | (m - 1ull);
    return m != 0ull && ((t + 1ull) & t) == 0ull;
}

// rowcnt[0][b][y] = raw chunks of the row, rowcnt[1][b][y] = run chunks (0 when use_runs is off: every chunk is raw)
__global__ void __launch_bounds__(EXP_THREADS)
export_count_kernel(const uint32_t *__restrict__ mask, size_t mask_pitch_w, size_t mask_fstride_w,
                    int *__restrict__ rowcnt, size_t half, int w, int h, int cpr, int use_runs) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int y = blockIdx.x * EXP_WARPS + warp, b = blockIdx.y;
    if (y >= h) return;
    const uint32_t *mrow = mask + (size_t)b * mask_fstride_w + (size_t)y * mask_pitch_w;
    const int words = (w + 31) >> 5;
    const unsigned lastmask = (w & 31) ? ((1u << (w & 31)) - 1u) : 0xffffffffu;
    int n = 0, nr = 0;
    for (int c0 = 0; c0 < cpr; c0 += 32) {
        const unsigned long long m = exp_chunk_mask(mrow, c0 + lane, cpr, words, lastmask);
        const unsigned any = __ballot_sync(0xffffffffu, m != 0ull);
        const unsigned one = use_runs ? __ballot_sync(0xffffffffu, exp_single_run(m)) : 0u;
        n += __popc(any & ~one);
        nr += __popc(one);
    }
    if (lane == 0) {
        rowcnt[(size_t)b * h + y] = n;
        rowcnt[half + (size_t)b * h + y] = nr;
    }
}

// exclusive scans of the two row-count arrays of one frame (one CTA per frame, blockIdx.y = which array); the totals go
// to device and host copies
__global__ void __launch_bounds__(EXP_THREADS)
export_scan_kernel(int *__restrict__ rowcnt, size_t half, int *__restrict__ totals_dev, int *__restrict__ totals_host,
                   int *__restrict__ run_totals_host, int h) {
    __shared__ int part[EXP_THREADS];
    const int b = blockIdx.x, tid = threadIdx.x, which = blockIdx.y;
    int *rc = rowcnt + (which ? half : 0) + (size_t)b * h;
    const int per = (h + EXP_THREADS - 1) / EXP_THREADS;
    const int lo = min(tid * per, h), hi = min(lo + per, h);
    int sum = 0;
    for (int i = lo; i < hi; i++) sum += rc[i];
    part[tid] = sum;
    __syncthreads();
    // Hillis-Steele inclusive scan over the 256 partial sums
    for (int d = 1; d < EXP_THREADS; d <<= 1) {
        const int v = tid >= d ? part[tid - d] : 0;
        __syncthreads();
        part[tid] += v;
        __syncthreads();
    }
    int run = part[tid] - sum;
    if (tid == EXP_THREADS - 1) {
        if (which == 0) {
            if (totals_dev) totals_dev[b] = part[tid];
            if (totals_host) totals_host[b] = part[tid];
        } else if (run_totals_host) {
            run_totals_host[b] = part[tid];
        }
    }
    for (int i = lo; i < hi; i++) { const int v = rc[i]; rc[i] = run; run += v; }
}

// one warp per row: the raw chunks of the row go to ids[] / data[] at the row's offset (32 lanes x 8 bytes = one chunk),
// the run chunks to runs[] (one 16-byte store by the lane that owns the chunk)
__global__ void __launch_bounds__(EXP_THREADS)
export_write_kernel(const uint32_t *__restrict__ mask, size_t mask_pitch_w, size_t mask_fstride_w,
                    const int32_t *__restrict__ labels, size_t labels_pitch_e, size_t labels_fstride_e,
                    const int *__restrict__ rowoff, size_t half, int32_t *__restrict__ ids, int32_t *__restrict__ data,
                    int32_t *__restrict__ runs, int w, int h, int cpr, int cap) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int y = blockIdx.x * EXP_WARPS + warp, b = blockIdx.y;
    if (y >= h) return;
    const uint32_t *mrow = mask + (size_t)b * mask_fstride_w + (size_t)y * mask_pitch_w;
    const int32_t *lrow = labels + (size_t)b * labels_fstride_e + (size_t)y * labels_pitch_e;
    const int words = (w + 31) >> 5;
    const unsigned lastmask = (w & 31) ? ((1u << (w & 31)) - 1u) : 0xffffffffu;
    int at = rowoff[(size_t)b * h + y];
    int at_r = runs ? rowoff[half + (size_t)b * h + y] : 0;
    int32_t *fid = ids + (size_t)b * cap;
    int32_t *fdata = data + (size_t)b * cap * VA_CHUNK_E;
    int4 *fruns = runs ? reinterpret_cast<int4 *>(runs) + (size_t)b * cap : nullptr;
    for (int c0 = 0; c0 < cpr; c0 += 32) {
        const unsigned long long m = exp_chunk_mask(mrow, c0 + lane, cpr, words, lastmask);
        unsigned bal = __ballot_sync(0xffffffffu, m != 0ull);
        const unsigned one = runs ? __ballot_sync(0xffffffffu, exp_single_run(m)) : 0u;
        // run chunks: every lane that owns one writes its own record (their offsets follow from the ballot)
        if ((one >> lane) & 1u) {
            const int slot = at_r + __popc(one & ((1u << lane) - 1u));
            if (slot < cap) {
                const int x = (c0 + lane) * VA_CHUNK_E + (__ffsll((long long)m) - 1);
                fruns[slot] = make_int4(y * cpr + c0 + lane, lrow[x], (int)(unsigned)m, (int)(unsigned)(m >> 32));
            }
        }
        at_r += __popc(one);
        bal &= ~one;
        while (bal) {
            const int cc = c0 + __ffs(bal) - 1;
            bal &= bal - 1;
            if (at < cap) {
                if (lane == 0) fid[at] = y * cpr + cc;
                // labels beyond the row's width are not part of the image: they travel as zeros
                const int x = cc * VA_CHUNK_E + 2 * lane;
                int2 v = make_int2(0, 0);
                if (x + 1 < w) v = *reinterpret_cast<const int2 *>(lrow + x);
                else if (x < w) v.x = lrow[x];
                *reinterpret_cast<int2 *>(fdata + (size_t)at * VA_CHUNK_E + 2 * lane) = v;
            }
            at++;
        }
    }
}

extern "C" int va_label_export_chunks(va_ctx *ctx, va_stream stream,
                                      const uint32_t *mask, size_t mask_pitch_w, size_t mask_fstride_w,
                                      const int32_t *labels, size_t labels_pitch_e, size_t labels_fstride_e,
                                      int w, int h, int batch,
                                      int32_t *ids, int32_t *data, int32_t *n_chunks, int32_t *n_chunks_dev,
                                      int32_t *runs, int32_t *n_runs, int cap) {
    VA_CHECK_CTX(ctx);
    VA_REQUIRE(ctx, mask && n_chunks && ((ids && data && labels) || (!ids && !data)), "va_label_export_chunks: null pointer");
    VA_REQUIRE(ctx, (!runs || (n_runs && data)) && (!(n_runs && data) || runs),
               "va_label_export_chunks: runs needs n_runs and the raw export; n_runs alone only with the count-only call");
    VA_REQUIRE(ctx, w > 0 && h > 0 && batch > 0 && batch <= 65535 && cap > 0, "va_label_export_chunks: bad size");
    VA_REQUIRE(ctx, !data || (labels_pitch_e >= (size_t)w && labels_pitch_e % 2 == 0 && labels_fstride_e % 2 == 0 && va_aligned(labels, 8)),
               "va_label_export_chunks: label rows must be 8-byte aligned");
    VA_REQUIRE(ctx, !runs || va_aligned(runs, 16), "va_label_export_chunks: runs must be 16-byte aligned");
    VA_REQUIRE(ctx, mask_pitch_w >= (size_t)((w + 31) / 32), "va_label_export_chunks: pitch smaller than a row");
    if (w > ctx->max_w || h > ctx->max_h || batch > ctx->max_batch)
        VA_FAIL(ctx, VA_ERR_CAPACITY, "va_label_export_chunks: %dx%dx%d exceeds the ctx capacity %dx%dx%d", w, h, batch,
                ctx->max_w, ctx->max_h, ctx->max_batch);
    const size_t half = (size_t)ctx->max_h * ctx->max_batch;          // raw counts, then run counts
    if (!ctx->exp_rowoff) {
        if (cudaMalloc((void **)&ctx->exp_rowoff, 2 * half * sizeof(int)) != cudaSuccess) {
            cudaGetLastError();
            VA_FAIL(ctx, VA_ERR_NOMEM, "va_label_export_chunks: cannot allocate the row offsets");
        }
    }
    const int cpr = va_div_up(w, VA_CHUNK_E);
    VA_REQUIRE(ctx, va_scratch_acquire(ctx, stream, 3) == 0, "va_label_export_chunks: cannot order the scratch");
    const dim3 grid(va_div_up(h, EXP_WARPS), batch);
    { auto k = export_count_kernel;
      VA_LAUNCH(ctx, k, grid, EXP_THREADS, 0, stream, mask, mask_pitch_w, mask_fstride_w, ctx->exp_rowoff, half, w, h, cpr, n_runs ? 1 : 0); }
    { auto k = export_scan_kernel;
      const dim3 grid_s(batch, n_runs ? 2 : 1);
      VA_LAUNCH(ctx, k, grid_s, EXP_THREADS, 0, stream, ctx->exp_rowoff, half, n_chunks_dev, n_chunks, n_runs, h); }
    if (data) {
      auto k = export_write_kernel;
      VA_LAUNCH(ctx, k, grid, EXP_THREADS, 0, stream, mask, mask_pitch_w, mask_fstride_w, labels, labels_pitch_e, labels_fstride_e,
                (const int *)ctx->exp_rowoff, half, ids, data, runs, w, h, cpr, cap); }
    VA_REQUIRE(ctx, va_scratch_release(ctx, stream, 3) == 0, "va_label_export_chunks: cannot order the scratch");
    return VA_OK;
}

// ---------------------------------------------------------------------------------------------------------
// host side: dense (batch, h, pitch_e) int32 images from exported chunks.  `dirty_ids` / `n_dirty` describe which chunks
// of `dense` are non-zero from its previous use (in: cleared here; out: the chunks written now).  Plain C++ threads, no
// CUDA call: this is the host half of the transfer and runs while the device works on the next blocks.
// ---------------------------------------------------------------------------------------------------------
extern "C" int va_host_densify_chunks(int32_t *dense, size_t pitch_e, size_t fstride_e, int w, int h, int batch,
                                      const int32_t *ids, const int32_t *data, const int32_t *n_chunks,
                                      const int32_t *runs, const int32_t *n_runs, int cap,
                                      int32_t *dirty_ids, int32_t *n_dirty, int threads) {
    if (!dense || !ids || !data || !n_chunks || !dirty_ids || !n_dirty || w <= 0 || h <= 0 || batch <= 0 || cap <= 0 ||
        (runs == nullptr) != (n_runs == nullptr))
        return VA_ERR_INVALID;
    const int cpr = (w + VA_CHUNK_E - 1) / VA_CHUNK_E;
    const int last_e = w - (cpr - 1) * VA_CHUNK_E;          // elements of the last chunk of a row
    for (int b = 0; b < batch; b++) {
        const long long nr = n_runs ? n_runs[b] : 0;
        if (n_chunks[b] < 0 || n_chunks[b] > cap || nr < 0 || nr > cap || n_chunks[b] + nr > cap || n_dirty[b] < 0 || n_dirty[b] > cap)
            return VA_ERR_CAPACITY;
    }
    auto frame = [&](int b) {
        int32_t *img = dense + (size_t)b * fstride_e;
        int32_t *dirty = dirty_ids + (size_t)b * cap;
        for (int i = 0; i < n_dirty[b]; i++) {
            const int id = dirty[i], y = id / cpr, c = id - y * cpr;
            std::memset(img + (size_t)y * pitch_e + (size_t)c * VA_CHUNK_E, 0, sizeof(int32_t) * (c == cpr - 1 ? last_e : VA_CHUNK_E));
        }
        const int32_t *fid = ids + (size_t)b * cap;
        const int32_t *fdata = data + (size_t)b * cap * VA_CHUNK_E;
        const int n = n_chunks[b];
        for (int i = 0; i < n; i++) {
            const int id = fid[i], y = id / cpr, c = id - y * cpr;
            std::memcpy(img + (size_t)y * pitch_e + (size_t)c * VA_CHUNK_E, fdata + (size_t)i * VA_CHUNK_E,
                        sizeof(int32_t) * (c == cpr - 1 ? last_e : VA_CHUNK_E));
        }
        std::memcpy(dirty, fid, sizeof(int32_t) * (size_t)n);
        // run chunks: (id, label, low / high word of the pixel mask); the set bits are one run
        const int nr = n_runs ? n_runs[b] : 0;
        const int32_t *fr = runs ? runs + (size_t)b * cap * 4 : nullptr;
        for (int i = 0; i < nr; i++) {
            const int id = fr[4 * i], lab = fr[4 * i + 1], y = id / cpr, c = id - y * cpr;
            const unsigned long long m = ((unsigned long long)(uint32_t)fr[4 * i + 3] << 32) | (uint32_t)fr[4 * i + 2];
            dirty[n + i] = id;
            if (m == 0ull) continue;                         // never written by the device; nothing to fill
            const int start = __builtin_ctzll(m), len = __builtin_popcountll(m);
            const int room = (c == cpr - 1 ? last_e : VA_CHUNK_E) - start;
            int32_t *p = img + (size_t)y * pitch_e + (size_t)c * VA_CHUNK_E + start;
            for (int k = 0; k < (len < room ? len : room); k++) p[k] = lab;
        }
        n_dirty[b] = n + nr;
    };
    if (threads <= 1 || batch == 1) {
        for (int b = 0; b < batch; b++) frame(b);
        return VA_OK;
    }
    if (threads > batch) threads = batch;
    std::vector<std::thread> pool;
    pool.reserve(threads);
    for (int t = 0; t < threads; t++)
        pool.emplace_back([&, t]() {
            for (int b = t; b < batch; b += threads) frame(b);
        });
    for (auto &th : pool) th.join();
    return VA_OK;
}
