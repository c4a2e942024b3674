// va_label.cu -- K5: connected-component labelling of packed bit masks, replacing
// scipy.ndimage.label at video/analysis/regions.py:162 (4-connectivity by default,
// labels 1..n numbered in raster order of each component's first pixel), plus the
// per-label areas of regions.py:165-169.
//
// Run-based union-find.  The unit of equivalence is a horizontal run of foreground
// pixels, represented by the index (y << LOG) + x of its first pixel; the forest
// `parent` lives in ctx scratch and is only ever touched at run starts, so its HBM
// traffic scales with the number of runs, not pixels.  Hooking always points the larger
// index at the smaller one, hence the root of a component is its first pixel in raster
// order, and ranking the roots in index order IS the canonical numbering.
//
// Kernels A and C walk the mask with four LANES per image row (most rows are empty: a few
// instructions per word).  Kernels B and G own one row per warp (lane = 32-pixel word); run
// starts that continue across words are resolved with one ballot + one shuffle per 32 words.
// (A lane-per-row merge was measured slower: rows of one blob then hook in lock step and the
// concurrent finds walk the chain that is being built.)
//   A init     parent[s] = s for every run start s
//   B merge    for every maximal overlap segment between a run in row y and one in row
//              y-1: union(run start, run start)        (atomicMin hooking)
//   C rank     the k-th root of a row (k = 1, 2, ... in x order) gets parent = -k; roots per
//              row are counted (no pointer chasing here: a root is a run start that is still
//              its own parent)
//   D scan     exclusive prefix of the per-row root counts (one CTA per frame) -> n
//   G write    label(run) = rowoff[row of its root] + k; one walk to the root per run, the
//              labels of a 1024-pixel chunk are assembled in a per-warp shared-memory slab
//              and leave as 16-byte stores (512 contiguous bytes per warp instruction)
// Algorithmic HBM bytes per frame: N/8 (mask) + 4N (labels).
#include "va_device.cuh"

#define LAB_THREADS 256
#define LAB_WARPS (LAB_THREADS / 32)
#define FULL 0xffffffffu

// ---- row helpers ---------------------------------------------------------------------
__device__ __forceinline__ unsigned lab_load_word(const uint32_t *row, int wpw, int j, unsigned lastmask) {
    if (j >= wpw) return 0u;
    unsigned v = row[j];
    if (j == wpw - 1) v &= lastmask;
    return v;
}

// For the 32 words of a chunk (lane = word): x position of the start of the run that
// reaches bit 0 of this word from the left (== 32 * word index when nothing continues),
// and `top` = start of the run that reaches past bit 31 (== 32 * (word + 1) when none).
// `carry` is `top` of the word preceding the chunk.
__device__ __forceinline__ void lab_scan_chunk(unsigned wd, int lane, int base_word, int carry, int &st_in, int &top) {
    const unsigned nf = __ballot_sync(FULL, wd != FULL);
    const unsigned below = nf & ((1u << lane) - 1u);
    const int k = 31 - __clz((int)below);                        // nearest not-all-ones word to the left, -1: none
    const unsigned wk = __shfl_sync(FULL, wd, k < 0 ? 0 : k);
    st_in = k < 0 ? carry : 32 * (base_word + k) + (32 - __clz((int)~wk));
    top = wd != FULL ? 32 * (base_word + lane) + (32 - __clz((int)~wd)) : st_in;
}

// start (x position) of the run containing bit `bit` of word `wd` (bit must be set)
__device__ __forceinline__ int lab_run_start(unsigned wd, int bit, int word_index, int st_in) {
    const unsigned z = ~wd & ((1u << bit) - 1u);
    return z ? 32 * word_index + (32 - __clz((int)z)) : st_in;
}

// all words of a row (up to 4096 pixels) are requested before the first one is used, so a
// warp pays the memory latency once per row instead of once per 1024-pixel chunk
#define LAB_PREFETCH(pre, row)                                                          \
    unsigned pre[4];                                                                    \
    _Pragma("unroll") for (int c_ = 0; c_ < 4; c_++) pre[c_] = lab_load_word(row, wpw, 32 * c_ + lane, lastmask)
#define LAB_PICK(pre, base, row)                                                                   \
    ((base) == 0 ? pre[0] : (base) == 32 ? pre[1] : (base) == 64 ? pre[2] : (base) == 96 ? pre[3] \
                 : lab_load_word(row, wpw, (base) + lane, lastmask))

// first pixels of the runs of a word; `pbit` = bit 31 of the word to the left
__device__ __forceinline__ unsigned lab_run_starts(unsigned wd, unsigned pbit) {
    return wd & ~((wd << 1) | pbit);
}

// ---- union-find ------------------------------------------------------------------------
// a node is a root while parent == self; kernel C later marks roots with negative values
__device__ __forceinline__ int lab_find(const int *parent, int x) {
    int p;
    while ((p = va_ld_cg(parent + x)) != x && p >= 0) x = p;
    return x;
}
// find with path halving, used while trees are still being hooked (kernel B).  Rows are
// merged concurrently in no particular order, which would otherwise grow chains as long as a
// component is tall.  Halving only rewrites non-root nodes with one of their ancestors
// (parent values only ever decrease), and a hook is only trusted when atomicMin saw a root,
// so the plain store cannot lose an equivalence.
__device__ __forceinline__ int lab_find_halve(int *parent, int x) {
    int p = va_ld_cg(parent + x);
    while (p != x) {
        const int gp = va_ld_cg(parent + p);
        if (gp != p) parent[x] = gp;
        x = p;
        p = gp;
    }
    return x;
}
__device__ __forceinline__ void lab_union(int *parent, int a, int b) {
    while (true) {
        a = lab_find_halve(parent, a);
        b = lab_find_halve(parent, b);
        if (a == b) return;
        if (a < b) { const int t = a; a = b; b = t; }
        const int old = atomicMin(parent + a, b);
        if (old == a) return;        // a was a root and now hangs under b
        a = old;                     // lost a race: keep merging what a pointed to with b
    }
}

// =====================================================================================
// Kernels A and C walk the mask with four lanes per image row (8 rows per warp): each lane owns a
// quarter of the row's words, fetches them 16 at a time with 16-byte loads that are all in flight
// together, and scans them sequentially.  Most rows of a real mask are empty, and with a lane per
// quarter row an empty stretch costs a few instructions per word instead of a whole warp; run
// continuation across words is a carried bit instead of a warp scan.
// =====================================================================================
#define LAB_GROUP 4            // words per 16-byte load
#define LAB_BATCH 4            // loads in flight per lane

// words [j, j + 16) of a row (0 beyond the part / row end, last word masked)
__device__ __forceinline__ void lab_load_batch(const uint32_t *row, int j, int jend, int wpw, unsigned lastmask, bool vec,
                                               bool valid, unsigned (&wd)[LAB_GROUP * LAB_BATCH]) {
#pragma unroll
    for (int g = 0; g < LAB_BATCH; g++) {
        const int jg = j + LAB_GROUP * g;
        if (valid && vec && jg + LAB_GROUP <= jend) {
            const uint4 q = *reinterpret_cast<const uint4 *>(row + jg);
            wd[4 * g] = q.x; wd[4 * g + 1] = q.y; wd[4 * g + 2] = q.z; wd[4 * g + 3] = q.w;
        } else {
#pragma unroll
            for (int k = 0; k < LAB_GROUP; k++) wd[4 * g + k] = (valid && jg + k < jend) ? row[jg + k] : 0u;
        }
    }
    const int last = wpw - 1 - j;
#pragma unroll
    for (int k = 0; k < LAB_GROUP * LAB_BATCH; k++)
        if (k == last) wd[k] &= lastmask;
}

// ---- A: init -----------------------------------------------------------------------------
__global__ void __launch_bounds__(LAB_THREADS)
label_init_kernel(const uint32_t *__restrict__ mask, size_t mpw, size_t mfw,
                  int *__restrict__ parent, int LOG, size_t pf, int *__restrict__ rowflag, int w, int h, int batch, int vec) {
    const int lane = threadIdx.x & 31;
    const int wpw = (w + 31) >> 5;
    const unsigned lastmask = (w & 31) ? ((1u << (w & 31)) - 1u) : FULL;
    const int ppw = ((wpw + 15) >> 4) << 2;           // words per quarter row, a multiple of 4
    const int wpf = (h + 7) >> 3;                     // warps per frame
    const int total = wpf * batch;
    for (int wi = blockIdx.x * LAB_WARPS + (threadIdx.x >> 5); wi < total; wi += gridDim.x * LAB_WARPS) {
        const int b = wi / wpf, y = (wi - b * wpf) * 8 + (lane >> 2);
        const int j0 = (lane & 3) * ppw, jend = min(j0 + ppw, wpw);
        const bool valid = y < h && j0 < wpw;
        const uint32_t *row = mask + (size_t)b * mfw + (size_t)(y < h ? y : 0) * mpw;
        int *pr = parent + (size_t)b * pf + ((size_t)y << LOG);
        unsigned prev_top = (valid && j0 > 0) ? (row[j0 - 1] >> 31) : 0u;
        unsigned any = 0;
        for (int j = j0; j < jend; j += LAB_GROUP * LAB_BATCH) {
            unsigned wd[LAB_GROUP * LAB_BATCH];
            lab_load_batch(row, j, jend, wpw, lastmask, vec != 0, valid, wd);
#pragma unroll
            for (int k = 0; k < LAB_GROUP * LAB_BATCH; k++) {
                unsigned starts = lab_run_starts(wd[k], prev_top);
                any |= wd[k];
                while (starts) {
                    const int bit = __ffs((int)starts) - 1;
                    starts &= starts - 1;
                    const int x = 32 * (j + k) + bit;
                    pr[x] = (y << LOG) + x;
                }
                prev_top = wd[k] >> 31;
            }
        }
        // does the row have foreground at all?  (kernel B skips the others without touching the mask)
        any |= __shfl_xor_sync(FULL, any, 1);
        any |= __shfl_xor_sync(FULL, any, 2);
        if ((lane & 3) == 0 && y < h) rowflag[(size_t)b * h + y] = any != 0u;
    }
}

// ---- B: merge with the row above -----------------------------------------------------------
// one "probe" row: the row above shifted by dx in {-1, 0, +1} (dx != 0 only for 8-connectivity)
__device__ __forceinline__ void lab_merge_probe(int *pr, unsigned cur, int cur_stin, unsigned pu, int pu_stin,
                                                unsigned ov_prev_bit, int lane, int base, int rowc, int rowu,
                                                int dx, bool up_bit0) {
    const unsigned ov = cur & pu;
    unsigned ss = ov & ~((ov << 1) | ov_prev_bit);     // first pixel of every overlap segment
    while (ss) {
        const int bit = __ffs((int)ss) - 1;
        ss &= ss - 1;
        const int a = lab_run_start(cur, bit, base + lane, cur_stin);
        int bq = lab_run_start(pu, bit, base + lane, pu_stin);      // in the shifted row
        // undo the shift: probe(x) = up(x - dx)
        if (dx == 1) bq -= 1;
        else if (dx == -1) bq += (bq == 0 && up_bit0) ? 0 : 1;
        lab_union(pr, rowc + a, rowu + bq);
    }
}

__global__ void __launch_bounds__(LAB_THREADS, 8)
label_merge_kernel(const uint32_t *__restrict__ mask, size_t mpw, size_t mfw,
                   int *__restrict__ parent, int LOG, size_t pf, const int *__restrict__ rowflag,
                   int w, int h, int batch, int conn8, int R) {
    // grid = (groups of 8 R rows, frames): R consecutive rows per warp (R > 1 for narrow frames, where a row is too
    // little work for a warp's start-up; see lab_rows_per_warp)
    const int lane = threadIdx.x & 31;
    const int wpw = (w + 31) >> 5;
    const unsigned lastmask = (w & 31) ? ((1u << (w & 31)) - 1u) : FULL;
    const int yw = (blockIdx.x * LAB_WARPS + (threadIdx.x >> 5)) * R;
    // flags of rows yw - 1 .. yw + R - 1 (kernel A): lane i holds the flag of row yw - 1 + i
    const int yf = yw - 1 + lane;
    const int myflag = (lane <= R && yf >= 0 && yf < h) ? rowflag[(size_t)blockIdx.y * h + yf] : 0;
    const unsigned flags = __ballot_sync(FULL, myflag != 0);
    for (int rr = 0; rr < R; rr++) {
        const int b = blockIdx.y, y = yw + rr;
        if (y >= h) break;
        // a row can only be merged with the one above if both have foreground at all
        if (y == 0 || ((flags >> rr) & 3u) != 3u) continue;
        const uint32_t *mc = mask + (size_t)b * mfw + (size_t)y * mpw;
        const uint32_t *mu = mc - mpw;
        int *pr = parent + (size_t)b * pf;
        const int rowc = y << LOG, rowu = (y - 1) << LOG;
        bool up_bit0 = false;
        int carry_c = 0, carry_u = 0, carry_l = 0, carry_r = 0;
        unsigned ovp0 = 0, ovpl = 0, ovpr = 0;          // bit 31 of the overlap word before the chunk
        unsigned up_prev_top = 0;                        // bit 31 of the up word before the chunk
        LAB_PREFETCH(prec, mc);
        LAB_PREFETCH(preu, mu);
        for (int base = 0; base < wpw; base += 32) {
            const unsigned cur = LAB_PICK(prec, base, mc);
            const unsigned up = LAB_PICK(preu, base, mu);
            if (base == 0) up_bit0 = (__shfl_sync(FULL, up, 0) & 1u) != 0;
            // nothing can be merged in this chunk unless both rows have foreground in or next to it
            // (8-connectivity: the first pixel of the next chunk of the row above also counts)
            const unsigned nxt = conn8 ? __shfl_sync(FULL, LAB_PICK(preu, base + 32, mu), 0) : 0u;
            const bool any_cur = __any_sync(FULL, cur != 0u);
            const bool any_up = __any_sync(FULL, up != 0u) || up_prev_top || (nxt & 1u);
            if (!any_cur || !any_up) {
                // no overlap segment starts here; carries follow from the words alone
                int cs, ct, us, ut;
                if (any_cur) { lab_scan_chunk(cur, lane, base, carry_c, cs, ct); carry_c = __shfl_sync(FULL, ct, 31); }
                else carry_c = 32 * (base + 32);
                if (__any_sync(FULL, up != 0u)) {
                    lab_scan_chunk(up, lane, base, carry_u, us, ut);
                    carry_u = __shfl_sync(FULL, ut, 31);
                    if (conn8) {
                        const unsigned upw = __shfl_up_sync(FULL, up, 1);
                        const unsigned upl = (up << 1) | (lane ? (upw >> 31) : up_prev_top);
                        const unsigned upn = __shfl_down_sync(FULL, up, 1);
                        const unsigned upr = (up >> 1) | ((lane < 31 ? (upn & 1u) : (nxt & 1u)) << 31);
                        int ls, lt, rs, rt;
                        lab_scan_chunk(upl, lane, base, carry_l, ls, lt);
                        lab_scan_chunk(upr, lane, base, carry_r, rs, rt);
                        carry_l = __shfl_sync(FULL, lt, 31);
                        carry_r = __shfl_sync(FULL, rt, 31);
                    }
                } else {
                    carry_u = 32 * (base + 32);
                    if (conn8) {
                        // shifted rows: only the bits leaking in from the neighbouring chunks can be set
                        const unsigned upl = lane == 0 ? up_prev_top : 0u;
                        const unsigned upr = lane == 31 ? ((nxt & 1u) << 31) : 0u;
                        int ls, lt, rs, rt;
                        lab_scan_chunk(upl, lane, base, carry_l, ls, lt);
                        lab_scan_chunk(upr, lane, base, carry_r, rs, rt);
                        carry_l = __shfl_sync(FULL, lt, 31);
                        carry_r = __shfl_sync(FULL, rt, 31);
                    }
                }
                ovp0 = ovpl = ovpr = 0;
                up_prev_top = __shfl_sync(FULL, up, 31) >> 31;
                continue;
            }
            int cs, ct, us, ut;
            lab_scan_chunk(cur, lane, base, carry_c, cs, ct);
            lab_scan_chunk(up, lane, base, carry_u, us, ut);
            {
                const unsigned ov = cur & up;
                const unsigned pv = __shfl_up_sync(FULL, ov, 1);
                lab_merge_probe(pr, cur, cs, up, us, lane ? (pv >> 31) : ovp0, lane, base, rowc, rowu, 0, up_bit0);
                ovp0 = __shfl_sync(FULL, ov, 31) >> 31;
            }
            if (conn8) {
                // upl(x) = up(x - 1)
                const unsigned upw = __shfl_up_sync(FULL, up, 1);
                const unsigned upl = (up << 1) | (lane ? (upw >> 31) : up_prev_top);
                // upr(x) = up(x + 1): needs bit 0 of the next word (next chunk for lane 31)
                const unsigned upn = __shfl_down_sync(FULL, up, 1);
                const unsigned upr = (up >> 1) | ((lane < 31 ? (upn & 1u) : (nxt & 1u)) << 31);
                int ls, lt, rs, rt;
                lab_scan_chunk(upl, lane, base, carry_l, ls, lt);
                lab_scan_chunk(upr, lane, base, carry_r, rs, rt);
                {
                    const unsigned ov = cur & upl;
                    const unsigned pv = __shfl_up_sync(FULL, ov, 1);
                    lab_merge_probe(pr, cur, cs, upl, ls, lane ? (pv >> 31) : ovpl, lane, base, rowc, rowu, 1, up_bit0);
                    ovpl = __shfl_sync(FULL, ov, 31) >> 31;
                }
                {
                    const unsigned ov = cur & upr;
                    const unsigned pv = __shfl_up_sync(FULL, ov, 1);
                    lab_merge_probe(pr, cur, cs, upr, rs, lane ? (pv >> 31) : ovpr, lane, base, rowc, rowu, -1, up_bit0);
                    ovpr = __shfl_sync(FULL, ov, 31) >> 31;
                }
                carry_l = __shfl_sync(FULL, lt, 31);
                carry_r = __shfl_sync(FULL, rt, 31);
            }
            up_prev_top = __shfl_sync(FULL, up, 31) >> 31;
            carry_c = __shfl_sync(FULL, ct, 31);
            carry_u = __shfl_sync(FULL, ut, 31);
        }
    }
}

// ---- C: rank the roots inside their row, count them -------------------------------------------
// A root is a run start that is still its own parent after all merges.  The k-th root of a row
// in x order gets parent = -k (no pointer chasing here).  Four lanes per row: every lane counts the
// roots of its quarter, a 4-lane prefix gives its first rank, a second sweep assigns.
__global__ void __launch_bounds__(LAB_THREADS)
label_flatten_kernel(const uint32_t *__restrict__ mask, size_t mpw, size_t mfw,
                     int *__restrict__ parent, int LOG, size_t pf, int *__restrict__ rowcnt,
                     int w, int h, int batch, int vec) {
    const int lane = threadIdx.x & 31;
    const int wpw = (w + 31) >> 5;
    const unsigned lastmask = (w & 31) ? ((1u << (w & 31)) - 1u) : FULL;
    const int ppw = ((wpw + 15) >> 4) << 2;
    const int wpf = (h + 7) >> 3;
    const int total = wpf * batch;
    for (int wi = blockIdx.x * LAB_WARPS + (threadIdx.x >> 5); wi < total; wi += gridDim.x * LAB_WARPS) {
        const int b = wi / wpf, y = (wi - b * wpf) * 8 + (lane >> 2);
        const int j0 = (lane & 3) * ppw, jend = min(j0 + ppw, wpw);
        const bool valid = y < h && j0 < wpw;
        const uint32_t *row = mask + (size_t)b * mfw + (size_t)(y < h ? y : 0) * mpw;
        int *pr = parent + (size_t)b * pf + ((size_t)y << LOG);
        const unsigned top0 = (valid && j0 > 0) ? (row[j0 - 1] >> 31) : 0u;
        int rank = 0;
        for (int sweep = 0; sweep < 2; sweep++) {
            unsigned prev_top = top0;
            int nroots = 0;
            for (int j = j0; j < jend; j += LAB_GROUP * LAB_BATCH) {
                unsigned wd[LAB_GROUP * LAB_BATCH];
                lab_load_batch(row, j, jend, wpw, lastmask, vec != 0, valid, wd);
#pragma unroll
                for (int k = 0; k < LAB_GROUP * LAB_BATCH; k++) {
                    unsigned starts = lab_run_starts(wd[k], prev_top);
                    while (starts) {
                        const int bit = __ffs((int)starts) - 1;
                        starts &= starts - 1;
                        const int x = 32 * (j + k) + bit;
                        if (pr[x] == (y << LOG) + x) {
                            nroots++;
                            if (sweep) pr[x] = -(rank + nroots);
                        }
                    }
                    prev_top = wd[k] >> 31;
                }
            }
            if (sweep) break;
            // inclusive prefix over the four lanes of the row
            int inc = nroots;
            const int q = lane & 3;
            int v = __shfl_up_sync(FULL, inc, 1); if (q >= 1) inc += v;
            v = __shfl_up_sync(FULL, inc, 2); if (q >= 2) inc += v;
            rank = inc - nroots;
            if (q == 3 && y < h) rowcnt[(size_t)b * h + y] = inc;
            if (!__any_sync(FULL, nroots != 0)) break;         // nothing to rank in these 8 rows
        }
    }
}

// ---- D: exclusive scan of row counts, one CTA per frame --------------------------------------
__global__ void __launch_bounds__(LAB_THREADS)
label_scan_kernel(int *__restrict__ rowcnt, int *__restrict__ counts, int h) {
    const int b = blockIdx.x, tid = threadIdx.x;
    int *rc = rowcnt + (size_t)b * h;
    const int per = (h + LAB_THREADS - 1) / LAB_THREADS;
    const int lo = min(tid * per, h), hi = min(lo + per, h);
    int sum = 0;
    for (int i = lo; i < hi; i++) sum += rc[i];
    // exclusive prefix of the 256 partial sums: shuffle scan inside each warp, then over the 8 warp totals
    const int lane = tid & 31, wid = tid >> 5;
    int inc = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(FULL, inc, d);
        if (lane >= d) inc += v;
    }
    __shared__ int wtot[LAB_WARPS];
    if (lane == 31) wtot[wid] = inc;
    __syncthreads();
    int wbase = 0, total = 0;
#pragma unroll
    for (int i = 0; i < LAB_WARPS; i++) {
        const int v = wtot[i];
        if (i < wid) wbase += v;
        total += v;
    }
    if (tid == 0 && counts) counts[b] = total;
    int run = wbase + inc - sum;
    for (int i = lo; i < hi; i++) { const int v = rc[i]; rc[i] = run; run += v; }
}

// ---- G: write the label image --------------------------------------------------------------------
// Two steps per 1024-pixel chunk of a row, both inside one warp (no block barrier):
//   fill   lane = word: for each run of the word fetch its label once and write it to the
//          run's pixels in a per-warp shared-memory slab
//   store  lane = 4 consecutive pixels: background pixels become 0, foreground pixels read
//          the slab; one 16-byte store per lane, 512 contiguous bytes per warp instruction
// The slab index is skewed by one word per 32 pixels so that both steps are conflict-free.
#define LAB_SKEW(p) ((p) + ((p) >> 5))
// four consecutive labels of a row starting at pixel x (x % 4 == 0): one 16-byte store for int32 labels, one 8-byte
// store for int16 labels (scipy.ndimage.label(..., output=np.int16): half the bytes to write and to copy to the host)
__device__ __forceinline__ void lab_store4(int32_t *lr, int x, int w, int4 v, bool vec) {
    if (vec && x + 4 <= w) {
        *reinterpret_cast<int4 *>(lr + x) = v;
    } else {
        if (x < w) lr[x] = v.x;
        if (x + 1 < w) lr[x + 1] = v.y;
        if (x + 2 < w) lr[x + 2] = v.z;
        if (x + 3 < w) lr[x + 3] = v.w;
    }
}
__device__ __forceinline__ void lab_store4(int16_t *lr, int x, int w, int4 v, bool vec) {
    if (vec && x + 4 <= w) {
        *reinterpret_cast<uint2 *>(lr + x) = make_uint2(((unsigned)v.x & 0xffffu) | ((unsigned)v.y << 16),
                                                        ((unsigned)v.z & 0xffffu) | ((unsigned)v.w << 16));
    } else {
        if (x < w) lr[x] = (int16_t)v.x;
        if (x + 1 < w) lr[x + 1] = (int16_t)v.y;
        if (x + 2 < w) lr[x + 2] = (int16_t)v.z;
        if (x + 3 < w) lr[x + 3] = (int16_t)v.w;
    }
}

template <typename T>
__global__ void __launch_bounds__(LAB_THREADS)
label_write_kernel(const uint32_t *__restrict__ mask, size_t mpw, size_t mfw,
                   const int *__restrict__ parent, int LOG, size_t pf, const int *__restrict__ rowoff,
                   const int *__restrict__ rowflag,
                   T *__restrict__ labels, size_t lpe, size_t lfe, int w, int h, int batch, int vec_out, int R) {
    // grid = (groups of 8 R rows, frames): R consecutive rows per warp (R > 1 for narrow frames)
    // per-warp slab of min(w, 1024) pixels (+ skew): narrow frames leave room for more CTAs per SM, and the walk to the
    // roots is a chain of dependent loads that only more rows in flight can hide
    VA_DYN_SMEM(int, slab_all);
    const int lane = threadIdx.x & 31;
    const int slab_px = (w < 1024 ? ((w + 31) & ~31) : 1024);
    int *slab = slab_all + (threadIdx.x >> 5) * (slab_px + (slab_px >> 5));
    const int wpw = (w + 31) >> 5;
    const unsigned lastmask = (w & 31) ? ((1u << (w & 31)) - 1u) : FULL;
    const int yw = (blockIdx.x * LAB_WARPS + (threadIdx.x >> 5)) * R;
    const int myflag = (lane < R && yw + lane < h) ? rowflag[(size_t)blockIdx.y * h + yw + lane] : 0;
    const unsigned flags = __ballot_sync(FULL, myflag != 0);      // bit i: row yw + i has foreground (flag of kernel A)
    for (int rr = 0; rr < R; rr++) {
        const int b = blockIdx.y, y = yw + rr;
        if (y >= h) break;
        T *lr = labels + (size_t)b * lfe + (size_t)y * lpe;
        if (!((flags >> rr) & 1u)) {
            // background row: zeros straight to memory, the mask is not even read
            const int4 z = make_int4(0, 0, 0, 0);
            for (int x = 4 * lane; x < w; x += 128) lab_store4(lr, x, w, z, vec_out != 0);
            continue;
        }
        const uint32_t *mr = mask + (size_t)b * mfw + (size_t)y * mpw;
        const int *pf_ = parent + (size_t)b * pf;
        const int *pr = pf_ + ((size_t)y << LOG);
        const int *ro = rowoff + (size_t)b * h;
        int carry = 0;
        LAB_PREFETCH(pre, mr);
        for (int base = 0; base < wpw; base += 32) {
            const unsigned wd = LAB_PICK(pre, base, mr);
            const int nwords = min(32, wpw - base);
            if (!__any_sync(FULL, wd != 0u)) {
                // background only: zeros straight to memory
                const int4 z = make_int4(0, 0, 0, 0);
                for (int it = 0; 4 * it < nwords; it++) lab_store4(lr, 32 * base + 128 * it + 4 * lane, w, z, vec_out != 0);
                carry = 32 * (base + 32);
                continue;
            }
            int st_in, top;
            lab_scan_chunk(wd, lane, base, carry, st_in, top);
            // ---- fill: one lookup per run.  A word that holds a single run (the usual case inside a blob) keeps its label in
            // a register and hands it to the store step by shuffle; only words with several runs go through the slab
            // (writing a 32-pixel run to the slab pixel by pixel was 23 % of the kernel's instructions on blob masks)
            auto run_label = [&](int start_x) {
                // walk to the root: every root holds -k (kernel C), nothing else is negative; the
                // forest is final, so ordinary cached loads are fine and chains are short (halving)
                int idx = (y << LOG) + start_x;
                int p = pr[start_x];
                while (p >= 0) { idx = p; p = pf_[idx]; }
                return ro[idx >> LOG] - p;                         // p == -k
            };
            int single = -1;                                       // labels are >= 1
            const unsigned heads = wd & ~(wd << 1);                // first pixels of the runs of this word (bit 0: may continue)
            if (wd != 0u && (heads & (heads - 1u)) == 0u) {
                const int bit = __ffs((int)wd) - 1;
                single = run_label(bit == 0 ? st_in : 32 * (base + lane) + bit);
            } else {
                unsigned rem = wd;
                while (rem) {
                    const int bit = __ffs((int)rem) - 1;
                    const unsigned t = ~(wd >> bit);
                    const int ones = t ? __ffs((int)t) - 1 : 32;
                    const int lab = run_label(bit == 0 ? st_in : 32 * (base + lane) + bit);
                    const int p0 = 32 * lane + bit;
                    for (int k = 0; k < ones; k++) slab[LAB_SKEW(p0 + k)] = lab;
                    rem &= ~((ones >= 32 ? FULL : ((1u << ones) - 1u)) << bit);
                }
            }
            __syncwarp();
            // ---- store: 128 pixels per step
            for (int it = 0; 4 * it < nwords; it++) {
                const int src = 4 * it + (lane >> 3);
                const unsigned wi = __shfl_sync(FULL, wd, src);
                const int sl = __shfl_sync(FULL, single, src);
                const unsigned nib = (wi >> (4 * (lane & 7))) & 0xFu;
                const int px = 128 * it + 4 * lane;
                const int x = 32 * base + px;
                int4 v = make_int4(0, 0, 0, 0);
                if (nib) {
                    if (sl >= 0) {
                        v = make_int4((nib & 1u) ? sl : 0, (nib & 2u) ? sl : 0, (nib & 4u) ? sl : 0, (nib & 8u) ? sl : 0);
                    } else {
                        if (nib & 1u) v.x = slab[LAB_SKEW(px)];
                        if (nib & 2u) v.y = slab[LAB_SKEW(px + 1)];
                        if (nib & 4u) v.z = slab[LAB_SKEW(px + 2)];
                        if (nib & 8u) v.w = slab[LAB_SKEW(px + 3)];
                    }
                }
                lab_store4(lr, x, w, v, vec_out != 0);
            }
            __syncwarp();
            carry = __shfl_sync(FULL, top, 31);
        }
    }
}

// rows a warp of the row-per-warp kernels takes one after the other.  Measured (blob masks): the merge gains from 2 rows
// at 1080p (forest 0.110 -> 0.101 ms per 128 frames) and 3 at VGA (0.085 -> 0.077 per 256 frames: a 640-pixel row is 20
// mask words -- too little per warp start-up); the write gains only on narrow frames (VGA 0.103 -> 0.094) and loses on
// wide ones (1080p 0.194 -> 0.203 with 2 rows)
static int lab_rows_per_warp(int w, bool merge) {
    if (getenv("VA_LABEL_ROWS")) { const int r = atoi(getenv("VA_LABEL_ROWS")); return r < 1 ? 1 : r > 8 ? 8 : r; }   // tuning only
    if (!merge) return w <= 1024 ? 2 : 1;
    const int r = 2048 / (w > 0 ? w : 1);
    return r < 2 ? 2 : r > 8 ? 8 : r;
}

// phases A-D: afterwards every root holds -(rank in its row), rowcnt holds the exclusive prefix of
// the per-row root counts and counts[b] the number of components
// scratch set `slot` (0: allocated by va_create, 1: on first use) -- two sets let the label write of one batch
// run beside the forest kernels of the next
static int label_scratch(va_ctx *ctx, const char *name, int slot, int **parent, int **rowcnt) {
    VA_REQUIRE(ctx, slot == 0 || slot == 1, "%s: scratch slot must be 0 or 1", name);
    if (slot == 1 && !ctx->lab_parent1) {
        const size_t n_parent = ctx->lab_pitch * (size_t)ctx->max_h * (size_t)ctx->max_batch;
        if (cudaMalloc((void **)&ctx->lab_parent1, n_parent * sizeof(int32_t)) != cudaSuccess ||
            cudaMalloc((void **)&ctx->lab_rowcnt1, (size_t)2 * ctx->max_h * ctx->max_batch * sizeof(int32_t)) != cudaSuccess) {
            cudaGetLastError();
            cudaFree(ctx->lab_parent1);
            ctx->lab_parent1 = nullptr;
            VA_FAIL(ctx, VA_ERR_NOMEM, "%s: cannot allocate the second labelling scratch", name);
        }
    }
    *parent = slot ? ctx->lab_parent1 : ctx->lab_parent;
    *rowcnt = slot ? ctx->lab_rowcnt1 : ctx->lab_rowcnt;
    return VA_OK;
}

static int label_forest(va_ctx *ctx, va_stream stream, const char *name,
                        const uint32_t *mask, size_t mask_pitch_w, size_t mask_fstride_w,
                        int32_t *counts, int w, int h, int batch, int connectivity, int *LOG_out, size_t *pf_out,
                        int slot = 0) {
    VA_REQUIRE(ctx, w > 0 && h > 0 && batch > 0, "%s: bad size", name);
    VA_REQUIRE(ctx, connectivity == 4 || connectivity == 8, "%s: connectivity must be 4 or 8", name);
    VA_REQUIRE(ctx, mask_pitch_w >= (size_t)((w + 31) / 32), "%s: pitch smaller than a row", name);
    if (w > ctx->max_w || h > ctx->max_h || batch > ctx->max_batch)
        VA_FAIL(ctx, VA_ERR_CAPACITY, "%s: %dx%dx%d exceeds the ctx capacity %dx%dx%d", name, w, h, batch,
                ctx->max_w, ctx->max_h, ctx->max_batch);
    int LOG = 5;
    while (((size_t)1 << LOG) < ctx->lab_pitch) LOG++;
    const size_t pf = ctx->lab_pitch * (size_t)ctx->max_h;
    int *parent, *rowcnt;
    { const int rc = label_scratch(ctx, name, slot, &parent, &rowcnt); if (rc != VA_OK) return rc; }
    int *rowflag = rowcnt + (size_t)ctx->max_h * ctx->max_batch;               // does the row have foreground?
    VA_REQUIRE(ctx, va_scratch_acquire(ctx, stream, slot) == 0, "%s: cannot order the scratch set", name);
    const int vec = va_aligned(mask, 16) && mask_pitch_w % 4 == 0 && mask_fstride_w % 4 == 0;
    const int grid_a = va_div_up((long long)((h + 7) / 8) * batch, LAB_WARPS);          // four-lanes-per-row kernels
    { auto k = label_init_kernel;
      VA_LAUNCH(ctx, k, grid_a, LAB_THREADS, 0, stream, mask, mask_pitch_w, mask_fstride_w, parent, LOG, pf, rowflag, w, h, batch, vec); }
    { auto k = label_merge_kernel;
      VA_REQUIRE(ctx, batch <= 65535, "%s: more than 65535 frames in one call", name);
      const int R = lab_rows_per_warp(w, true);
      const dim3 grid_b(va_div_up(h, LAB_WARPS * R), batch);
      VA_LAUNCH(ctx, k, grid_b, LAB_THREADS, 0, stream, mask, mask_pitch_w, mask_fstride_w, parent, LOG, pf, (const int *)rowflag,
                w, h, batch, connectivity == 8 ? 1 : 0, R); }
    { auto k = label_flatten_kernel;
      VA_LAUNCH(ctx, k, grid_a, LAB_THREADS, 0, stream, mask, mask_pitch_w, mask_fstride_w, parent, LOG, pf, rowcnt, w, h, batch, vec); }
    { auto k = label_scan_kernel;
      VA_LAUNCH(ctx, k, batch, LAB_THREADS, 0, stream, rowcnt, counts, h); }
    VA_REQUIRE(ctx, va_scratch_release(ctx, stream, slot) == 0, "%s: cannot order the scratch set", name);
    *LOG_out = LOG;
    *pf_out = pf;
    return VA_OK;
}

template <typename T>
static int label_write(va_ctx *ctx, va_stream stream, const char *name,
                       const uint32_t *mask, size_t mask_pitch_w, size_t mask_fstride_w,
                       T *labels, size_t labels_pitch_e, size_t labels_fstride_e,
                       int w, int h, int batch, int LOG, size_t pf, int slot) {
    int *parent, *rowcnt;
    { const int rc = label_scratch(ctx, name, slot, &parent, &rowcnt); if (rc != VA_OK) return rc; }
    VA_REQUIRE(ctx, va_scratch_acquire(ctx, stream, slot) == 0, "%s: cannot order the scratch set", name);
    auto k = label_write_kernel<T>;
    const int vec_out = va_aligned(labels, 4 * sizeof(T)) && labels_pitch_e % 4 == 0 && labels_fstride_e % 4 == 0;
    const int R = lab_rows_per_warp(w, false);
    const dim3 grid(va_div_up(h, LAB_WARPS * R), batch);
    const int slab_px = (w < 1024 ? ((w + 31) & ~31) : 1024);
    const size_t smem = (size_t)LAB_WARPS * (slab_px + (slab_px >> 5)) * sizeof(int);
    VA_LAUNCH(ctx, k, grid, LAB_THREADS, smem, stream, mask, mask_pitch_w, mask_fstride_w, (const int *)parent, LOG, pf,
              (const int *)rowcnt, (const int *)(rowcnt + (size_t)ctx->max_h * ctx->max_batch),
              labels, labels_pitch_e, labels_fstride_e, w, h, batch, vec_out, R);
    VA_REQUIRE(ctx, va_scratch_release(ctx, stream, slot) == 0, "%s: cannot order the scratch set", name);
    return VA_OK;
}

extern "C" int va_label_bits(va_ctx *ctx, va_stream stream,
                             const uint32_t *mask, size_t mask_pitch_w, size_t mask_fstride_w,
                             int32_t *labels, size_t labels_pitch_e, size_t labels_fstride_e,
                             int32_t *counts, int w, int h, int batch, int connectivity) {
    VA_CHECK_CTX(ctx);
    VA_REQUIRE(ctx, mask && labels, "va_label_bits: null pointer");
    VA_REQUIRE(ctx, w <= 0 || labels_pitch_e >= (size_t)w, "va_label_bits: pitch smaller than a row");
    int LOG;
    size_t pf;
    const int rc = label_forest(ctx, stream, "va_label_bits", mask, mask_pitch_w, mask_fstride_w, counts, w, h, batch,
                                connectivity, &LOG, &pf);
    if (rc != VA_OK) return rc;
    return label_write(ctx, stream, "va_label_bits", mask, mask_pitch_w, mask_fstride_w, labels, labels_pitch_e,
                       labels_fstride_e, w, h, batch, LOG, pf, 0);
}

// va_label_bits in two halves, each on the stream it is given, sharing scratch set `slot` (0 or 1): the union-find
// forest of a batch (counts[] is complete after it) and the write of its label image.  With two scratch sets the
// write of batch k (pure stores, HBM-bound) can run on one stream while the forest of batch k + 1 (dependent loads,
// latency-bound) runs on another; the caller orders write(k) after forest(k) and forest(k + 2) after write(k).
extern "C" int va_label_forest(va_ctx *ctx, va_stream stream,
                               const uint32_t *mask, size_t mask_pitch_w, size_t mask_fstride_w,
                               int32_t *counts, int w, int h, int batch, int connectivity, int slot) {
    VA_CHECK_CTX(ctx);
    VA_REQUIRE(ctx, mask && counts, "va_label_forest: null pointer");
    int LOG;
    size_t pf;
    return label_forest(ctx, stream, "va_label_forest", mask, mask_pitch_w, mask_fstride_w, counts, w, h, batch,
                        connectivity, &LOG, &pf, slot);
}

// the same write with int16 labels (scipy.ndimage.label(..., output=np.int16)); the caller checks counts[] <= 32767
// (scipy raises "insufficient bit-depth in requested output type" otherwise; labels above 32767 would wrap here)
extern "C" int va_label_write_i16(va_ctx *ctx, va_stream stream,
                                  const uint32_t *mask, size_t mask_pitch_w, size_t mask_fstride_w,
                                  int16_t *labels, size_t labels_pitch_e, size_t labels_fstride_e,
                                  int w, int h, int batch, int slot) {
    VA_CHECK_CTX(ctx);
    VA_REQUIRE(ctx, mask && labels, "va_label_write_i16: null pointer");
    VA_REQUIRE(ctx, w > 0 && h > 0 && batch > 0 && batch <= 65535, "va_label_write_i16: bad size");
    VA_REQUIRE(ctx, labels_pitch_e >= (size_t)w && mask_pitch_w >= (size_t)((w + 31) / 32), "va_label_write_i16: pitch smaller than a row");
    if (w > ctx->max_w || h > ctx->max_h || batch > ctx->max_batch)
        VA_FAIL(ctx, VA_ERR_CAPACITY, "va_label_write_i16: %dx%dx%d exceeds the ctx capacity %dx%dx%d", w, h, batch,
                ctx->max_w, ctx->max_h, ctx->max_batch);
    int LOG = 5;
    while (((size_t)1 << LOG) < ctx->lab_pitch) LOG++;
    return label_write(ctx, stream, "va_label_write_i16", mask, mask_pitch_w, mask_fstride_w, labels, labels_pitch_e,
                       labels_fstride_e, w, h, batch, LOG, ctx->lab_pitch * (size_t)ctx->max_h, slot);
}

extern "C" int va_label_write(va_ctx *ctx, va_stream stream,
                              const uint32_t *mask, size_t mask_pitch_w, size_t mask_fstride_w,
                              int32_t *labels, size_t labels_pitch_e, size_t labels_fstride_e,
                              int w, int h, int batch, int slot) {
    VA_CHECK_CTX(ctx);
    VA_REQUIRE(ctx, mask && labels, "va_label_write: null pointer");
    VA_REQUIRE(ctx, w > 0 && h > 0 && batch > 0 && batch <= 65535, "va_label_write: bad size");
    VA_REQUIRE(ctx, labels_pitch_e >= (size_t)w && mask_pitch_w >= (size_t)((w + 31) / 32), "va_label_write: pitch smaller than a row");
    if (w > ctx->max_w || h > ctx->max_h || batch > ctx->max_batch)
        VA_FAIL(ctx, VA_ERR_CAPACITY, "va_label_write: %dx%dx%d exceeds the ctx capacity %dx%dx%d", w, h, batch,
                ctx->max_w, ctx->max_h, ctx->max_batch);
    int LOG = 5;
    while (((size_t)1 << LOG) < ctx->lab_pitch) LOG++;
    return label_write(ctx, stream, "va_label_write", mask, mask_pitch_w, mask_fstride_w, labels, labels_pitch_e,
                       labels_fstride_e, w, h, batch, LOG, ctx->lab_pitch * (size_t)ctx->max_h, slot);
}

// =====================================================================================
// per-region statistics straight from the mask: the label image is never written.
// Row-run form of the raw moments of cv2.moments(region.astype(uint8)) (video/analysis/image.py:350)
// and the bounding box of regions.py:113-149, exact in 64-bit integers:
//   a run x0..x1 of row y adds  n = x1-x0+1,  sx = n (x0+x1) / 2,  sxx = sum x^2,
//   m00 += n  m10 += sx  m01 += n y  m20 += sxx  m11 += y sx  m02 += n y^2
// =====================================================================================
__global__ void __launch_bounds__(LAB_THREADS)
region_stats_init_kernel(long long *__restrict__ stats, int max_regions, const int *__restrict__ counts) {
    const int b = blockIdx.y;
    const int n = min(counts[b], max_regions);
    for (int l = blockIdx.x * LAB_THREADS + threadIdx.x; l < n; l += gridDim.x * LAB_THREADS) {
        long long *s = stats + ((size_t)b * max_regions + l) * VA_REGION_FIELDS;
#pragma unroll
        for (int k = 0; k < 6; k++) s[k] = 0;
        s[6] = s[7] = 0x7fffffff;        // xmin, ymin
        s[8] = s[9] = -1;                // xmax, ymax
    }
}

__global__ void __launch_bounds__(LAB_THREADS)
region_stats_kernel(const uint32_t *__restrict__ mask, size_t mpw, size_t mfw,
                    const int *__restrict__ parent, int LOG, size_t pf, const int *__restrict__ rowoff,
                    const int *__restrict__ rowflag,
                    long long *__restrict__ stats, int max_regions, int w, int h, int batch) {
    // grid = (groups of 8 rows, frames): one row per warp; background rows (flag of kernel A) have nothing to add
    const int lane = threadIdx.x & 31;
    const int wpw = (w + 31) >> 5;
    const unsigned lastmask = (w & 31) ? ((1u << (w & 31)) - 1u) : FULL;
    {
        const int b = blockIdx.y, y = blockIdx.x * LAB_WARPS + (threadIdx.x >> 5);
        if (y >= h || rowflag[(size_t)b * h + y] == 0) return;
        const uint32_t *mr = mask + (size_t)b * mfw + (size_t)y * mpw;
        const int *pf_ = parent + (size_t)b * pf;
        const int *pr = pf_ + ((size_t)y << LOG);
        const int *ro = rowoff + (size_t)b * h;
        long long *sb = stats + (size_t)b * max_regions * VA_REGION_FIELDS;
        int carry = 0;
        LAB_PREFETCH(pre, mr);
        for (int base = 0; base < wpw; base += 32) {
            const unsigned wd = LAB_PICK(pre, base, mr);
            if (!__any_sync(FULL, wd != 0u)) { carry = 32 * (base + 32); continue; }
            int st_in, top;
            lab_scan_chunk(wd, lane, base, carry, st_in, top);
            // the part of every run that lies in this word is one contribution
            unsigned rem = wd;
            while (rem) {
                const int bit = __ffs((int)rem) - 1;
                const unsigned t = ~(wd >> bit);
                const int ones = t ? __ffs((int)t) - 1 : 32;
                const int start_x = bit == 0 ? st_in : 32 * (base + lane) + bit;
                int idx = (y << LOG) + start_x;
                int p = pr[start_x];
                while (p >= 0) { idx = p; p = pf_[idx]; }
                const int lab = ro[idx >> LOG] - p;                // 1 .. n
                rem &= ~((ones >= 32 ? FULL : ((1u << ones) - 1u)) << bit);
                if (lab > max_regions) continue;
                const long long x0 = 32 * (base + lane) + bit, x1 = x0 + ones - 1, n = ones;
                const long long sx = n * (x0 + x1) / 2;
                // sum_{x=x0}^{x1} x^2 = F(x1) - F(x0 - 1), F(k) = k (k + 1) (2k + 1) / 6
                const long long f1 = x1 * (x1 + 1) * (2 * x1 + 1) / 6, f0 = (x0 - 1) * x0 * (2 * x0 - 1) / 6;
                unsigned long long *s = reinterpret_cast<unsigned long long *>(sb + (size_t)(lab - 1) * VA_REGION_FIELDS);
                atomicAdd(s + 0, (unsigned long long)n);
                atomicAdd(s + 1, (unsigned long long)sx);
                atomicAdd(s + 2, (unsigned long long)(n * y));
                atomicAdd(s + 3, (unsigned long long)(f1 - f0));
                atomicAdd(s + 4, (unsigned long long)(sx * y));
                atomicAdd(s + 5, (unsigned long long)(n * y * y));
                long long *sl = reinterpret_cast<long long *>(s);
                atomicMin(sl + 6, x0);
                atomicMin(sl + 7, (long long)y);
                atomicMax(sl + 8, x1);
                atomicMax(sl + 9, (long long)y);
            }
            carry = __shfl_sync(FULL, top, 31);
        }
    }
}

__global__ void __launch_bounds__(LAB_THREADS)
region_largest64_kernel(const long long *__restrict__ stats, int max_regions, const int *__restrict__ counts,
                        int *__restrict__ largest) {
    __shared__ long long best_a[LAB_THREADS];
    __shared__ int best_l[LAB_THREADS];
    const int b = blockIdx.x, tid = threadIdx.x;
    const int n = min(counts[b], max_regions);
    long long ba = 0;
    int bl = 0;
    for (int l = tid; l < n; l += LAB_THREADS) {
        const long long a = stats[((size_t)b * max_regions + l) * VA_REGION_FIELDS];
        if (a > ba) { ba = a; bl = l + 1; }
    }
    best_a[tid] = ba; best_l[tid] = bl;
    __syncthreads();
    // tree reduction; ties go to the smaller label (np.argmax returns the first maximum)
    for (int s = LAB_THREADS / 2; s > 0; s >>= 1) {
        if (tid < s) {
            const long long oa = best_a[tid + s];
            const int ol = best_l[tid + s];
            if (oa > best_a[tid] || (oa == best_a[tid] && oa > 0 && ol < best_l[tid])) { best_a[tid] = oa; best_l[tid] = ol; }
        }
        __syncthreads();
    }
    if (tid == 0) largest[b] = best_l[0];      // np.argmax(areas) + 1 (regions.py:169); 0 when the frame has no region
}

extern "C" int va_region_stats(va_ctx *ctx, va_stream stream,
                               const uint32_t *mask, size_t mask_pitch_w, size_t mask_fstride_w,
                               int64_t *stats, int max_regions, int32_t *counts, int32_t *largest,
                               int w, int h, int batch, int connectivity) {
    VA_CHECK_CTX(ctx);
    VA_REQUIRE(ctx, mask && stats && counts, "va_region_stats: null pointer");
    VA_REQUIRE(ctx, max_regions > 0, "va_region_stats: max_regions must be positive");
    int LOG;
    size_t pf;
    const int rc = label_forest(ctx, stream, "va_region_stats", mask, mask_pitch_w, mask_fstride_w, counts, w, h, batch,
                                connectivity, &LOG, &pf);
    if (rc != VA_OK) return rc;
    { auto k = region_stats_init_kernel;
      const dim3 grid(va_div_up(max_regions, LAB_THREADS) < 64 ? va_div_up(max_regions, LAB_THREADS) : 64, batch);
      VA_LAUNCH(ctx, k, grid, LAB_THREADS, 0, stream, (long long *)stats, max_regions, (const int *)counts); }
    { auto k = region_stats_kernel;
      const dim3 grid(va_div_up(h, LAB_WARPS), batch);
      VA_LAUNCH(ctx, k, grid, LAB_THREADS, 0, stream, mask, mask_pitch_w, mask_fstride_w, (const int *)ctx->lab_parent, LOG, pf,
                (const int *)ctx->lab_rowcnt, (const int *)(ctx->lab_rowcnt + (size_t)ctx->max_h * ctx->max_batch),
                (long long *)stats, max_regions, w, h, batch); }
    if (largest) {
        auto k = region_largest64_kernel;
        VA_LAUNCH(ctx, k, batch, LAB_THREADS, 0, stream, (const long long *)stats, max_regions, (const int *)counts, largest);
    }
    VA_REQUIRE(ctx, va_scratch_release(ctx, stream, 0) == 0, "va_region_stats: cannot order the scratch set");
    return VA_OK;
}

// =====================================================================================
// per-label areas + largest region (video/analysis/regions.py:165-169)
// =====================================================================================
__global__ void __launch_bounds__(LAB_THREADS)
region_area_kernel(const int32_t *__restrict__ labels, size_t lpe, size_t lfe,
                   int *__restrict__ areas, int max_labels, int w, int h, int batch) {
    const int lane = threadIdx.x & 31;
    const int chunks = (w + 31) >> 5;
    const unsigned total = (unsigned)((long long)chunks * h * batch);
    for (unsigned it = blockIdx.x * LAB_WARPS + (threadIdx.x >> 5); it < total;
         it += gridDim.x * LAB_WARPS) {
        const unsigned row = it / (unsigned)chunks;
        const int c = (int)(it - row * chunks);
        const int b = row / h, y = row - b * h;
        const int x = 32 * c + lane;
        const int lab = x < w ? labels[(size_t)b * lfe + (size_t)y * lpe + x] : 0;
        if (!__any_sync(FULL, lab != 0)) continue;
        // heads of runs of equal labels inside the 32-pixel chunk add the run length once
        const int left = __shfl_up_sync(FULL, lab, 1);
        const bool head = lane == 0 || left != lab;
        const unsigned heads = __ballot_sync(FULL, head);
        if (head && lab > 0 && lab <= max_labels) {
            const unsigned above = heads & ~((2u << lane) - 1u);       // heads to the right of me
            const int end = above ? __ffs((int)above) - 1 : 32;
            atomicAdd(areas + (size_t)b * max_labels + (lab - 1), end - lane);
        }
    }
}

__global__ void __launch_bounds__(LAB_THREADS)
region_largest_kernel(const int *__restrict__ areas, int max_labels, int *__restrict__ largest) {
    __shared__ int best_a[LAB_THREADS];
    __shared__ int best_l[LAB_THREADS];
    const int b = blockIdx.x, tid = threadIdx.x;
    int ba = 0, bl = 0;
    for (int l = tid; l < max_labels; l += LAB_THREADS) {
        const int a = areas[(size_t)b * max_labels + l];
        if (a > ba) { ba = a; bl = l + 1; }       // strided ascending: the first maximum wins below
    }
    best_a[tid] = ba; best_l[tid] = bl;
    __syncthreads();
    // tree reduction; ties go to the smaller label (np.argmax returns the first maximum)
    for (int s = LAB_THREADS / 2; s > 0; s >>= 1) {
        if (tid < s) {
            const int oa = best_a[tid + s], ol = best_l[tid + s];
            if (oa > best_a[tid] || (oa == best_a[tid] && oa > 0 && ol < best_l[tid])) { best_a[tid] = oa; best_l[tid] = ol; }
        }
        __syncthreads();
    }
    if (tid == 0) largest[b] = best_l[0];      // np.argmax(areas) + 1; 0 when the frame has no region
}

extern "C" int va_region_areas(va_ctx *ctx, va_stream stream,
                               const int32_t *labels, size_t labels_pitch_e, size_t labels_fstride_e,
                               int32_t *areas, int max_labels, int32_t *largest, int w, int h, int batch) {
    VA_CHECK_CTX(ctx);
    VA_REQUIRE(ctx, labels && areas, "va_region_areas: null pointer");
    VA_REQUIRE(ctx, w > 0 && h > 0 && batch > 0 && max_labels > 0, "va_region_areas: bad size");
    VA_CUDA(ctx, cudaMemsetAsync(areas, 0, (size_t)batch * max_labels * sizeof(int), (cudaStream_t)stream));
    const long long warps = (long long)((w + 31) / 32) * h * batch;
    const int grid = va_grid(ctx, (warps + LAB_WARPS - 1) / LAB_WARPS, 8);
    { auto k = region_area_kernel;
      VA_LAUNCH(ctx, k, grid, LAB_THREADS, 0, stream, labels, labels_pitch_e, labels_fstride_e, areas, max_labels, w, h, batch); }
    if (largest) {
        auto k = region_largest_kernel;
        VA_LAUNCH(ctx, k, batch, LAB_THREADS, 0, stream, (const int *)areas, max_labels, largest);
    }
    return VA_OK;
}
