// radix_sort.cu -- one-sweep LSD radix sort of 64-bit keys (+ optional 32-bit payload).
// See radix_sort.cuh for the role it plays on the path.
//
// Per pass and per 4096-key tile (256 threads x 16 keys, tiles taken by ticket):
//   1. keys are loaded warp-striped; each warp ranks its 512 keys digit by digit with
//      __match_any_sync into a warp-private shared histogram (stable);
//   2. thread d owns digit d: it turns the 8 warp histograms into offsets, publishes the tile's
//      digit count, and walks the predecessor tiles' published counts (decoupled look-back,
//      32-bit words: 2 status bits + 30 count bits) until it meets an inclusive prefix;
//   3. keys are placed in tile-sorted order in shared memory and written out in runs, so the
//      global stores of one digit are contiguous.
// Traffic: one read of the keys for all histograms, then one read + one write per executed pass.
#include "radix_sort.cuh"

namespace rt3d {

struct RsPassArgs {
    const unsigned long long *kin;
    unsigned long long *kout;
    const unsigned *vin;
    unsigned *vout;
    long long n;
    const int *n_dev;
    int total_bits;
    const int *total_bits_dev;
    int pass;
    const unsigned *hist;   // [8][256]
    unsigned *tile_state;   // this pass: [tiles][256]
    int *ticket;
    int *status;
};

__global__ void __launch_bounds__(kRsThreads) rs_hist_kernel(const unsigned long long *__restrict__ keys, long long n_host,
                                                             const int *n_dev, int total_bits_host,
                                                             const int *total_bits_dev, unsigned *hist) {
    __shared__ unsigned s_hist[kRsMaxPasses * 256];
    const long long n = n_dev ? (long long)*n_dev : n_host;
    const int total_bits = total_bits_dev ? *total_bits_dev : total_bits_host;
    const int npass = (total_bits + 7) / 8;
    for (int i = threadIdx.x; i < kRsMaxPasses * 256; i += kRsThreads) s_hist[i] = 0;
    __syncthreads();
    // warp-uniform trip count so that match_any always sees the whole warp
    const long long stride = (long long)gridDim.x * kRsThreads;
    for (long long base = (long long)blockIdx.x * kRsThreads; base < n; base += stride) {
        const long long i = base + threadIdx.x;
        const bool valid = i < n;
        const unsigned long long key = valid ? keys[i] : 0ull;
        rs_hist_add(s_hist, key, valid, npass);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < npass * 256; i += kRsThreads) {
        const unsigned c = s_hist[i];
        if (c) atomicAdd(&hist[i], c);
    }
}

template <bool HAS_VAL>
__global__ void __launch_bounds__(kRsThreads) rs_pass_kernel(const RsPassArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long *s_keys = reinterpret_cast<unsigned long long *>(smem_raw);              // kRsTile
    unsigned *s_vals = reinterpret_cast<unsigned *>(s_keys + kRsTile);                            // kRsTile (HAS_VAL)
    unsigned *s_whist = s_vals + (HAS_VAL ? kRsTile : 0);                                         // 8 * 256
    __shared__ unsigned s_gbase[256];
    __shared__ int s_dstart[256];
    __shared__ long long s_scat[256];
    __shared__ int s_tmp[8];
    __shared__ int s_tile;

    const long long n = a.n_dev ? (long long)*a.n_dev : a.n;
    const int total_bits = a.total_bits_dev ? *a.total_bits_dev : a.total_bits;
    const int shift = a.pass * 8;
    if (shift >= total_bits || n <= 0) return;  // pass not needed: data stays where it is
    const int num_tiles = (int)((n + kRsTile - 1) / kRsTile);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    {   // global exclusive digit offsets of this pass
        int tot;
        const int h = (int)a.hist[a.pass * 256 + tid];
        const int ex = block_excl_scan_256(h, s_tmp, tot);
        s_gbase[tid] = (unsigned)ex;
        __syncthreads();
    }

    for (;;) {
        if (tid == 0) s_tile = atomicAdd(a.ticket, 1);
        for (int i = tid; i < 8 * 256; i += kRsThreads) s_whist[i] = 0;
        __syncthreads();
        const int tile = s_tile;
        if (tile >= num_tiles) break;
        const long long tile_base = (long long)tile * kRsTile;
        const int tile_n = (int)((n - tile_base) < (long long)kRsTile ? (n - tile_base) : (long long)kRsTile);

        // ---- 1. load (warp-striped) and rank
        unsigned long long key[kRsItems];
        unsigned val[kRsItems];
        int rank[kRsItems];
        const long long wbase = tile_base + (long long)warp * (32 * kRsItems) + lane;
#pragma unroll
        for (int i = 0; i < kRsItems; ++i) {
            const long long idx = wbase + i * 32;
            key[i] = idx < n ? a.kin[idx] : ~0ull;  // padding sorts last and is never written
            if (HAS_VAL) val[i] = idx < n ? a.vin[idx] : 0u;
        }
        unsigned *wh = s_whist + warp * 256;
#pragma unroll
        for (int i = 0; i < kRsItems; ++i) {
            const unsigned d = (unsigned)((key[i] >> shift) & 255u);
            const unsigned peers = __match_any_sync(kFullMask, d);
            const int leader = __ffs(peers) - 1;
            unsigned pre = 0;
            if (lane == leader) {
                pre = wh[d];
                wh[d] = pre + (unsigned)__popc(peers);
            }
            pre = __shfl_sync(kFullMask, pre, leader);
            rank[i] = (int)(pre + (unsigned)__popc(peers & lanemask_lt()));
            __syncwarp();
        }
        __syncthreads();

        // ---- 2. thread d owns digit d
        {
            const int d = tid;
            unsigned run = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w) {
                const unsigned c = s_whist[w * 256 + d];
                s_whist[w * 256 + d] = run;
                run += c;
            }
            const int total = (int)run;  // includes the padding keys (digit 255 of the last tile)
            const unsigned real_total = (unsigned)(total - (d == 255 ? (kRsTile - tile_n) : 0));
            unsigned *st = a.tile_state + (size_t)tile * 256 + d;
            st_relaxed_u32(st, (tile == 0 ? kStPrefix32 : kStAgg32) | real_total);
            int tile_sum;
            const int dstart = block_excl_scan_256(total, s_tmp, tile_sum);
            s_dstart[d] = dstart;
            unsigned excl = 0;
            if (tile > 0) {
                for (int t = tile - 1;; --t) {
                    const unsigned *pt = a.tile_state + (size_t)t * 256 + d;
                    unsigned w;
                    int spins = 0;
                    do {
                        w = ld_relaxed_u32(pt);
                    } while ((w & kStMask32) == 0 && ++spins < kSpinLimit);
                    if ((w & kStMask32) == 0) { atomicOr(a.status, RT3D_DEV_LOOKBACK); break; }
                    excl += w & ~kStMask32;
                    if ((w & kStMask32) == kStPrefix32) break;
                }
                st_relaxed_u32(st, kStPrefix32 | (excl + real_total));
            }
            s_scat[d] = (long long)s_gbase[d] + (long long)excl - (long long)dstart;
        }
        __syncthreads();

        // ---- 3. tile-sorted order in shared memory, then run-contiguous global stores
#pragma unroll
        for (int i = 0; i < kRsItems; ++i) {
            const unsigned d = (unsigned)((key[i] >> shift) & 255u);
            const int pos = s_dstart[d] + (int)wh[d] + rank[i];
            s_keys[pos] = key[i];
            if (HAS_VAL) s_vals[pos] = val[i];
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < kRsItems; ++j) {
            const int p = tid + j * kRsThreads;
            if (p < tile_n) {
                const unsigned long long k = s_keys[p];
                const unsigned d = (unsigned)((k >> shift) & 255u);
                const long long g = s_scat[d] + p;
                a.kout[g] = k;
                if (HAS_VAL) a.vout[g] = s_vals[p];
            }
        }
        __syncthreads();
    }
}

static size_t rs_smem_bytes(bool has_val) {
    return (size_t)kRsTile * 8 + (has_val ? (size_t)kRsTile * 4 : 0) + 8 * 256 * 4;
}

int launch_sort(const SortArgs &a, cudaStream_t stream) {
    RT3D_REQUIRE(a.n >= 0 && a.n < (1ll << 30), RT3D_ERANGE, "sort: %lld keys outside [0, 2^30)", a.n);
    RT3D_REQUIRE(a.total_bits >= 0 && a.total_bits <= 64, RT3D_EINVAL, "sort: total_bits %d", a.total_bits);
    if (a.n == 0 || a.total_bits == 0) return RT3D_OK;
    const bool has_val = a.vals[0] != nullptr;
    const long long tiles = (a.n + kRsTile - 1) / kRsTile;
    int *tickets = (int *)a.state;
    unsigned *hist = sort_hist_ptr(a.state);
    unsigned *tile_state = (unsigned *)((char *)hist + align_up(kRsMaxPasses * 256 * sizeof(unsigned)));
    const int max_passes = sort_max_passes(a.total_bits);

    if (!a.hist_done) {
        const long long hb = (a.n + kRsThreads * 8 - 1) / (kRsThreads * 8);
        rs_hist_kernel<<<persistent_blocks(hb, 8), kRsThreads, 0, stream>>>(a.keys[0], a.n, a.n_dev, a.total_bits,
                                                                           a.total_bits_dev, hist);
        RT3D_LAUNCH_OK();
    }
    // the opt-in shared-memory size is a per-device function attribute: set it for the current device on every call
    // (a host-side table lookup; a process-wide "done" flag would leave a second GPU of the process at the 48 KB default)
    if (has_val)
        RT3D_CUDA_OK(cudaFuncSetAttribute(rs_pass_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rs_smem_bytes(true)));
    else
        RT3D_CUDA_OK(cudaFuncSetAttribute(rs_pass_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rs_smem_bytes(false)));
    const int grid = persistent_blocks(tiles, 3);
    for (int p = 0; p < max_passes; ++p) {
        RsPassArgs pa;
        pa.kin = a.keys[p & 1]; pa.kout = a.keys[(p + 1) & 1];
        pa.vin = a.vals[p & 1]; pa.vout = a.vals[(p + 1) & 1];
        pa.n = a.n; pa.n_dev = a.n_dev; pa.total_bits = a.total_bits; pa.total_bits_dev = a.total_bits_dev;
        pa.pass = p; pa.hist = hist;
        pa.tile_state = tile_state + (size_t)p * (size_t)(tiles + 1) * 256;
        pa.ticket = tickets + p; pa.status = a.status;
        if (has_val) rs_pass_kernel<true><<<grid, kRsThreads, rs_smem_bytes(true), stream>>>(pa);
        else rs_pass_kernel<false><<<grid, kRsThreads, rs_smem_bytes(false), stream>>>(pa);
        RT3D_LAUNCH_OK();
    }
    return RT3D_OK;
}

}  // namespace rt3d
