// fft_fast_dispatch.cu -- lookup and launch of the specialised kernels whose
// instantiations live in the generated fft_fast_table_*.cu files.
#include <stdint.h>
#include <stdlib.h>
#include "fft_fast.cuh"
#include "fft_warp.cuh"

namespace b2fast {

extern const FastEntry table_d_row[], table_d_col[], table_f_row[], table_f_col[];
extern const int table_d_row_count, table_d_col_count, table_f_row_count, table_f_col_count;

static int g_max_smem = 0;
static int g_sms = 148;

static const FastEntry *find(int prec, int n, int col, int code, int kind = -1, int persist = 0)
{
    const FastEntry *t, *any = nullptr;
    int cnt, i;
    if (prec == B2D_F64) { t = col ? table_d_col : table_d_row; cnt = col ? table_d_col_count : table_d_row_count; }
    else { t = col ? table_f_col : table_f_row; cnt = col ? table_f_col_count : table_f_row_count; }
    for (i = 0; i < cnt; ++i)
        if (t[i].n == n && t[i].code == code && t[i].persist == persist) {
            if (t[i].r2r_kind == kind) return &t[i];          // specialised for this kind (or the plain entry)
            if (t[i].r2r_kind < 0) any = &t[i];
        }
    return any;
}

void init(int max_smem)
{
    g_max_smem = max_smem;
    {
        int dev = 0, sms = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0) g_sms = sms;
    }
    b2warp::init(max_smem, g_sms);
    const FastEntry *tabs[4] = { table_d_row, table_d_col, table_f_row, table_f_col };
    const int cnts[4] = { table_d_row_count, table_d_col_count, table_f_row_count, table_f_col_count };
    for (int k = 0; k < 4; ++k)
        for (int i = 0; i < cnts[k]; ++i)
            if (tabs[k][i].smem > 48 * 1024)
                cudaFuncSetAttribute(tabs[k][i].func, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)tabs[k][i].smem);
}

// structural applicability (known at plan time).  kernel code = tile + 100*flavor + 1000*col;
// flavor 2 = COL kernel with the four-step twiddle fused in its store, flavor 3 = ROW load
// with COL (transposed) store.
// kernel code = tile + 100 * flavor (+ 1000 for COL kernels), flavors 0-9; codes 2000 + tile = flavor 10 (ROW)
// codes 2100 + tile = flavor 11 (COL: four-step first pass with the c2r merge on its load)
static inline int code_col(int code) { return (code >= 1000 && code < 2000) || (code >= 2100 && code < 2200); }
static inline int code_flavor(int code) { return code >= 2100 ? 11 : (code >= 2000 ? 10 : (code / 100) % 10); }

static const FastEntry *entry_for(const b2d_fft_pass &p)
{
    if (!p.kernel) return nullptr;
    const int col = code_col(p.kernel);
    const int flavor = code_flavor(p.kernel);
    if (flavor == 10) {
        // r2c split fused into the four-step's second pass: ROW load from dense scratch rows, split spectrum to
        // the user's interleaved output; n1 = bn[0] rows, a whole number of mirror-paired tiles
        if (p.pre_op || p.post_op != B2D_STORE_R2C_SPLIT || p.bluestein || p.npeer || p.load_col || !p.store_col) return nullptr;
        if (p.is != 2 || (p.os & 1) || (p.bos[0] & 1) || (p.bos[1] & 1) || (p.bos[2] & 1) || (p.bis[0] & 1) ||
            (p.bis[1] & 1) || (p.bis[2] & 1)) return nullptr;
        const FastEntry *e10 = find(p.prec, p.n, 0, p.kernel);
        if (!e10 || ((int)e10->smem > g_max_smem && g_max_smem)) return nullptr;
        if (p.bn[0] < e10->tpb || p.bn[0] % e10->tpb) return nullptr;
        return e10;
    }
    if (flavor == 11) {
        if (p.pre_op != B2D_LOAD_C2R_MERGE || p.post_op != B2D_STORE_TWIDDLE4 || p.bluestein || p.npeer ||
            !p.load_col || !p.store_col || !p.idx_mul || !p.aux2) return nullptr;
        if ((p.is & 1) || (p.os & 1) || p.bis[0] != 2 || p.bos[0] != 2) return nullptr;
        for (int i = 1; i < B2D_MAX_BATCH_DIMS; ++i)
            if ((p.bis[i] & 1) || (p.bos[i] & 1)) return nullptr;
        const FastEntry *e11 = find(p.prec, p.n, 1, p.kernel);
        if (e11 && (int)e11->smem > g_max_smem && g_max_smem) return nullptr;
        return e11;
    }
    if (flavor == 9) {
        // r2r kinds fused into the pass: real lines, so strides are in single reals
        if (p.pre_op != B2D_LOAD_R2R || p.post_op != B2D_STORE_R2R || p.bluestein || p.npeer) return nullptr;
        // ROW kernels read contiguous lines; their stores go through the line's own stride, so the lines may
        // also be stored transposed (row -> col)
        if ((p.load_col && !p.store_col) || col != p.load_col) return nullptr;
        const int64_t unit = p.r2r_pair ? 2 : 1;          // paired lines: batch dim 0 steps over two adjacent reals
        if (col ? (p.bis[0] != unit || p.bos[0] != unit || (p.r2r_pair && (p.pair_is != 1 || p.pair_os != 1)))
                : (p.is != 1 || (p.os != 1 && !p.store_col))) return nullptr;
        const FastEntry *e9 = find(p.prec, p.n, col, p.kernel, p.r2r_kind);
        if (e9 && (int)e9->smem > g_max_smem && g_max_smem) return nullptr;
        return e9;
    }
    if ((p.pre_op & B2D_LOAD_R2R) || (p.post_op & B2D_STORE_R2R)) return nullptr;
    if (flavor == 7) {
        if (!p.bluestein || p.pre_op != (B2D_LOAD_PAD | B2D_LOAD_CHIRP) ||
            p.post_op != (B2D_STORE_TRUNC | B2D_STORE_CHIRP_SCALE) || p.load_col || p.store_col || col || p.npeer)
            return nullptr;
    } else if (p.pre_op || p.bluestein) return nullptr;
    if (flavor == 7) { /* ops checked above */ }
    else if (flavor == 2) { if (p.post_op != B2D_STORE_TWIDDLE4 || !p.load_col || !p.store_col) return nullptr; }
    else if (flavor == 8) { if (p.post_op || !p.npeer || p.peer_rows <= 0 || !p.load_col || !p.store_col) return nullptr; }
    else if (p.post_op) return nullptr;
    if (flavor >= 4 && flavor <= 6 && !col) return nullptr;                          // L2-prefetch flavours: COL kernels only
    if ((flavor == 8) != (p.npeer && p.peer_rows > 0)) return nullptr;               // row-split stores: flavour 8 only
    if (flavor == 3) { if (p.load_col || !p.store_col || p.npeer) return nullptr; }
    else if (p.load_col != p.store_col || col != p.load_col) return nullptr;
    if ((p.is & 1) || (p.os & 1)) return nullptr;
    for (int i = 0; i < B2D_MAX_BATCH_DIMS; ++i)
        if ((p.bis[i] & 1) || (p.bos[i] & 1)) return nullptr;
    if (p.load_col && p.bis[0] != 2) return nullptr;                 // adjacent pencils on the load side
    if (p.store_col && p.bos[0] != 2) return nullptr;                // ... and on the store side
    if (!p.load_col && p.is != 2) return nullptr;                    // contiguous transforms
    if (!p.store_col && p.os != 2) return nullptr;
    const FastEntry *e = find(p.prec, p.n, col, p.kernel);
    if (e && (int)e->smem > g_max_smem && g_max_smem) return nullptr;
    return e;
}

int available(const b2d_fft_pass &p, int code)
{
    b2d_fft_pass q = p;
    q.kernel = code;
    if (code >= 3000) return b2warp::applicable(q, code);
    return entry_for(q) != nullptr;
}

size_t smem_bytes(const b2d_fft_pass &p)
{
    if (p.kernel >= 3000) return b2warp::smem_for(p, p.kernel);
    const FastEntry *e = entry_for(p);
    return e ? e->smem : 0;
}

int try_launch(const b2d_fft_pass &p, cudaStream_t st)
{
    if (p.kernel >= 3000) return b2warp::launch(p, st);
    const FastEntry *e = entry_for(p);
    if (!e) return 1;
    const size_t rs = p.prec == B2D_F32 ? 4 : 8;
    const bool reals = code_flavor(p.kernel) == 9;              // r2r: scalar accesses, no re/im pairing
    const intptr_t din = reals ? (intptr_t)rs : (const char *)p.in_im - (const char *)p.in_re;
    const intptr_t dout = reals ? (intptr_t)rs : (char *)p.out_im - (char *)p.out_re;   /* also tells the peer path whether to swap */
    // interleaved (im = re +- 1 scalar) and vector-aligned, else the generic kernel handles it
    if ((din != (intptr_t)rs && din != -(intptr_t)rs) || (dout != (intptr_t)rs && dout != -(intptr_t)rs)) return 1;
    const int swap_in = din < 0, swap_out = dout < 0;
    const uintptr_t lo_in = (uintptr_t)(swap_in ? p.in_im : p.in_re);
    uintptr_t lo_out = (uintptr_t)(swap_out ? p.out_im : p.out_re);
    if (p.npeer) {
        lo_out = 0;
        for (int i = 0; i < p.npeer; ++i) lo_out |= (uintptr_t)p.peer_out[i];
    }
    if (!reals && ((lo_in % (2 * rs)) || (lo_out % (2 * rs)))) return 1;
    int64_t tiles0 = (p.bn[0] + e->tpb - 1) / e->tpb;
    b2d_fft_pass q = p;
    q.tpb = e->tpb;                  // decode_block() uses the tile width
    if (code_flavor(p.kernel) == 11 && swap_in) return 1;     // the merge defines its own (swapped) load
    if (code_flavor(p.kernel) == 10) {
        if (swap_in || swap_out) return 1;
        // tiles of HALF = tpb / 2 rows + their mirrors over rows 1 .. n1/2, plus one tile for row 0
        q.aux_split = p.bn[0];
        tiles0 = p.bn[0] / e->tpb + 1;
        q.bn[0] = tiles0 * e->tpb;
    }
    const int64_t blocks = tiles0 * p.bn[1] * p.bn[2];
    if (blocks <= 0) return 0;
    if (blocks > 2147483647LL) return -1;
    int64_t grid = blocks;
    if (p.grid_limit > 0 && grid > p.grid_limit) {
        // an instantiation whose CTAs loop over the tiles, when the table has one for this kernel
        const FastEntry *pe = find(p.prec, p.n, code_col(p.kernel), p.kernel, -1, 1);
        if (pe && pe->tpb == e->tpb) { e = pe; grid = p.grid_limit; }
    }
    e->launch(q, swap_in, swap_out, (unsigned)grid, (long long)blocks, st);
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace b2fast
