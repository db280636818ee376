// stft_large.cuh -- transforms too long for one SM's shared memory (M = P*N > 16384, N = 8192 .. 65536).
//
// Same algorithm as stft_kernels.cuh -- P premultiplied sub-transforms of length N, in-place radix
// (R0,16,16,16) DIF with R0 = N/4096, digit-reversed output, conjugate-symmetry split, host-tabulated
// resample, threshold dB step -- but the first radix-R0 pass works global -> global, after which every
// 4096-point block is an independent transform that fits shared memory:
//
//   large_first   : PCM -> window -> W_M^(n q0) premultiply -> radix-R0 butterfly -> W_N^(lo q) -> Y   (global)
//   large_block   : one 4096-point block of Y: three radix-16 passes in shared memory, written back in
//                   natural order of the block's own bin index k (digit reversal undone in shared memory)
//   large_split   : bins kk = 1..M/2-1 from Z[kk], Z[M-kk] -> (L,R) magnitudes.  Bin kk = k*Q + c lives in
//                   block c at index k (Q = M/4096 blocks), i.e. Z is the transpose of the bin order: the kernel
//                   is a tiled transpose through shared memory, 512-byte runs on both the read and the write side
//   large_colour  : resample (lane per sample, taps read through L1/L2) + dB step + LUT -> RGBA
//
// Scratch per frame pair (mono) or frame (stereo): M float2 for Y/Z plus M/2 float2 of magnitudes.
#pragma once
#include "stft_kernels.cuh"

namespace sbx {

constexpr int LARGE_LOG2_BLOCK = 12;                       // 4096-point blocks
constexpr int LARGE_BLOCK = 1 << LARGE_LOG2_BLOCK;
constexpr int LARGE_THREADS = 256;

using G12 = Geom<LARGE_LOG2_BLOCK, 0, 2>;                  // the block transform: 4096 = 16 x 16 x 16

struct LargeArgs {
    KernelArgs k;               // tables as in the fused kernel; k.twp = planes of the FIRST (radix-R0) pass
    const float2 *twp_block;    // twiddle planes of the 4096-point block transform (G12 layout)
    float2 *zbuf;               // [units][M] transform scratch
    float2 *magbuf;             // [units][M/2] magnitudes by bin kk (index 0 unused)
    unsigned long long frame_base;   // first frame of this launch (chunked batches)
    uint32_t units;             // frame pairs (mono) or frames (stereo) in this launch
    uint32_t log2n, log2pad;
};

template <int LOG2N, int LOG2PAD> struct GeomL {
    static constexpr int N = 1 << LOG2N, PAD = 1 << LOG2PAD, LOG2M = LOG2N + LOG2PAD, M = 1 << LOG2M;
    static constexpr int LOG2R0 = LOG2N - LARGE_LOG2_BLOCK, R0 = 1 << LOG2R0;
    static_assert(LOG2R0 >= 1 && LOG2R0 <= 4, "first-pass radix 2..16");
    // bin kk = PAD*k + q0, k = d0 + R0*k_local  ->  position q0*N + d0*4096 + digitrev_4096(k_local)
    static __host__ __device__ constexpr int pos_of_bin(int kk)
    {
        const int q0 = kk & (PAD - 1), k = kk >> LOG2PAD;
        const int d0 = k & (R0 - 1), kl = k >> LOG2R0;
        return (q0 << LOG2N) + (d0 << LARGE_LOG2_BLOCK) + G12::pos_of_bin(kl);
    }
    // After large_block every 4096-point block holds its bins in natural order of the block-local index k, and
    // bin kk = k*Q + c sits in the block with "column" c = d0*PAD + q0, which is block number q0*R0 + d0.
    static constexpr int LOG2Q = LOG2M - LARGE_LOG2_BLOCK, Q = 1 << LOG2Q;   // blocks per unit
    static constexpr int TK = LARGE_BLOCK / Q;                                // k values per split tile (tile = 4096 bins)
    static constexpr int SPLIT_SMEM = 2 * TK * (Q + 1) * 8;                   // own + partner tiles, padded rows
    static __host__ __device__ constexpr int block_of_column(int c) { return ((c & (PAD - 1)) << LOG2R0) + (c >> LOG2PAD); }
};

// ---- pass 0: global PCM -> global Y --------------------------------------------------------------
// one thread per (unit, q0, lo): R0 inputs n = lo + 4096 r
template <int LOG2N, int LOG2PAD, int CH>
__device__ __forceinline__ void large_first(const LargeArgs &la, unsigned int block_id)
{
    using GL = GeomL<LOG2N, LOG2PAD>;
    constexpr int R = GL::R0;
    const KernelArgs &a = la.k;
    constexpr int TILES = LARGE_BLOCK / LARGE_THREADS;                 // thread tiles per (unit, q0)
    const unsigned int tile = block_id % TILES, q0 = (block_id / TILES) % GL::PAD, unit = block_id / (TILES * GL::PAD);
    const int lo = tile * LARGE_THREADS + threadIdx.x;
    const unsigned long long frame0 = la.frame_base + (unsigned long long)unit * (CH == 1 ? 2 : 1);
    const bool second_valid = CH == 1 && frame0 + 1 < a.n_frames;
    const float *pa = a.pcm + frame0 * (unsigned long long)a.hop * CH;
    const float *pb = pa + a.hop;
    float2 v[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int n = lo + r * LARGE_BLOCK;
        float2 s;
        if constexpr (CH == 2) {
            s = __ldg(reinterpret_cast<const float2 *>(pa) + n);
        } else {
            s.x = __ldg(pa + n);
            s.y = second_valid ? __ldg(pb + n) : 0.0f;
        }
        if (q0 == 0) {
            const float wn = __ldg(a.window + n);
            s.x *= wn;  s.y *= wn;
        } else {
            s = cmul(s, __ldg(a.wtw + (size_t)(q0 - 1) * GL::N + n));
        }
        v[r] = s;
    }
    Butterfly<R>::run(v);
    apply_twiddles<R>(v, a.twp + lo, LARGE_BLOCK);                     // W_N^(lo q)
    float2 *y = la.zbuf + (size_t)unit * GL::M + ((size_t)q0 << LOG2N) + lo;
#pragma unroll
    for (int q = 0; q < R; ++q) y[(size_t)q * LARGE_BLOCK] = v[q];
}

// ---- passes 1..3: one 4096-point block in shared memory -------------------------------------------
// phase 0: load + first radix-16 pass; phases 1,2: pass_inplace<G12,1|2>; phase 3: write back
__device__ __forceinline__ void large_block_load(float2 *z, const float2 *src, const float2 *twp_block)
{
    constexpr int S = 256;
    for (int w = threadIdx.x; w < LARGE_BLOCK / 16; w += LARGE_THREADS) {
        float2 v[16];
#pragma unroll
        for (int r = 0; r < 16; ++r) v[r] = src[w + r * S];
        Butterfly<16>::run(v);
        apply_twiddles<16>(v, twp_block + w, S);
        const int bx = G12::zsw(w);
#pragma unroll
        for (int q = 0; q < 16; ++q) z[SBX_ZAT(G12, bx, q * S)] = v[q];
    }
}
// write-back in natural order of the block-local bin index k (digit reversal undone here, in shared memory)
__device__ __forceinline__ void large_block_store(const float2 *z, float2 *dst)
{
    for (int k = threadIdx.x; k < LARGE_BLOCK; k += LARGE_THREADS) dst[k] = z[G12::zsw(G12::pos_of_bin(k))];
}

// ---- split: global Z -> global magnitudes (tiled transpose) -------------------------------------------
// One CTA = one tile of 4096 consecutive bins kk in [k0*Q, (k0+TK)*Q), k0 = tile*TK < 2048.
// phase 1 (large_split_load): for every column c, TK consecutive elements of block c (own) and the mirrored run
//   of block Q-c (partner, M-kk = (4095-k)*Q + (Q-c); column 0 mirrors onto itself at 4096-k) -> shared memory;
// phase 2 (large_split_emit): consecutive threads take consecutive kk.
template <int LOG2N, int LOG2PAD>
__device__ __forceinline__ void large_split_load(const LargeArgs &la, unsigned int block_id, float2 *own, float2 *par)
{
    using GL = GeomL<LOG2N, LOG2PAD>;
    const unsigned int unit = block_id / (GL::Q / 2), tile = block_id % (GL::Q / 2);
    const int k0 = tile * GL::TK;
    const float2 *zu = la.zbuf + (size_t)unit * GL::M;
#pragma unroll 8
    for (int i = threadIdx.x; i < LARGE_BLOCK; i += LARGE_THREADS) {
        const int c = i / GL::TK, j = i % GL::TK, k = k0 + j;
        own[j * (GL::Q + 1) + c] = zu[((size_t)GL::block_of_column(c) << LARGE_LOG2_BLOCK) + k];
        const int cp = (GL::Q - c) & (GL::Q - 1);
        const int kp = c != 0 ? LARGE_BLOCK - 1 - k : LARGE_BLOCK - k;          // k = 0, c = 0 is DC: no partner
        par[j * (GL::Q + 1) + c] = kp < LARGE_BLOCK ? zu[((size_t)GL::block_of_column(cp) << LARGE_LOG2_BLOCK) + kp]
                                                    : make_float2(0.0f, 0.0f);
    }
}

template <int LOG2N, int LOG2PAD, int CH>
__device__ __forceinline__ void large_split_emit(const LargeArgs &la, unsigned int block_id, const float2 *own, const float2 *par)
{
    using GL = GeomL<LOG2N, LOG2PAD>;
    const KernelArgs &a = la.k;
    const unsigned int unit = block_id / (GL::Q / 2), tile = block_id % (GL::Q / 2);
    const int kk0 = tile * LARGE_BLOCK;                                        // = k0 * Q
    const unsigned long long frame0 = la.frame_base + (unsigned long long)unit * (CH == 1 ? 2 : 1);
    for (int i = threadIdx.x; i < LARGE_BLOCK; i += LARGE_THREADS) {
        const int kk = kk0 + i;
        if (kk == 0) continue;                                                 // DC is dropped (fft.rs:81)
        const int j = i >> GL::LOG2Q, c = i & (GL::Q - 1);
        const float2 m = split_pair(own[j * (GL::Q + 1) + c], par[j * (GL::Q + 1) + c], a.scale);
        la.magbuf[(size_t)unit * (GL::M / 2) + kk] = m;
        if (a.mags_out != nullptr) {
            float2 *out = reinterpret_cast<float2 *>(a.mags_out) + frame0 * a.k_out + (kk - 1);
            if constexpr (CH == 1) {
                out[0] = make_float2(m.x, m.x);
                if (frame0 + 1 < a.n_frames) out[a.k_out] = make_float2(m.y, m.y);
            } else {
                out[0] = m;
            }
        }
    }
}

#ifndef SBX_EMU
template <int LOG2N, int LOG2PAD, int CH>
__global__ void __launch_bounds__(LARGE_THREADS) large_first_kernel(const LargeArgs la)
{
    large_first<LOG2N, LOG2PAD, CH>(la, blockIdx.x);
}

// grid = units * (M / 4096)
static __global__ void __launch_bounds__(LARGE_THREADS) large_block_kernel(const LargeArgs la)
{
    __shared__ __align__(16) float2 z[LARGE_BLOCK];
    float2 *blk = la.zbuf + (size_t)blockIdx.x * LARGE_BLOCK;
    KernelArgs a = la.k;
    a.twp = la.twp_block;
    large_block_load(z, blk, la.twp_block);
    __syncthreads();
    pass_inplace<G12, 1>(z, a);
    __syncthreads();
    pass_inplace<G12, 2>(z, a);
    __syncthreads();
    large_block_store(z, blk);
}

// grid = units * Q/2, dynamic shared memory = GeomL::SPLIT_SMEM
template <int LOG2N, int LOG2PAD, int CH>
__global__ void __launch_bounds__(LARGE_THREADS) large_split_kernel(const LargeArgs la)
{
    using GL = GeomL<LOG2N, LOG2PAD>;
    extern __shared__ __align__(16) unsigned char split_smem[];
    float2 *own = reinterpret_cast<float2 *>(split_smem);
    float2 *par = own + GL::TK * (GL::Q + 1);
    large_split_load<LOG2N, LOG2PAD>(la, blockIdx.x, own, par);
    __syncthreads();
    large_split_emit<LOG2N, LOG2PAD, CH>(la, blockIdx.x, own, par);
}

#endif  // SBX_EMU (the geometry of the cluster path is host-visible: tests/emu checks its column assignment)

// ---- cluster path: the whole transform of a unit in the distributed shared memory of one thread-block cluster ------
// M = 32768 .. 131072 (Q = M / 4096 = 8, 16, 32 blocks of 32 KB): a cluster of C = Q / 4 CTAs, four blocks (128 KB) each.
// The three trips through the global scratch of the four-kernel path (Y out, Y in / Z out, Z in) become
//   phase A  every CTA takes an equal share of the radix-R0 first-pass butterflies (PCM -> window -> premultiply ->
//            butterfly -> W_N^(lo q)) and SCATTERS each output into the block it belongs to -- a st.shared::cluster
//            into the owning CTA's shared memory, already at its swizzled slot;
//   -------- barrier.cluster --------
//   phase B  the four local 4096-point blocks: three in-place radix-16 passes each, four thread groups in lock step;
//   phase C  the L/R split, LOCAL: bin kk = k Q + c pairs with M - kk = (4095 - k) Q + (Q - c), and columns c and
//            Q - c are assigned to the same CTA (0 and Q / 2 pair with themselves), so no second exchange is needed;
//            magnitudes go to the global array large_colour_kernel reads.
// HBM scratch per unit: M/2 magnitudes written once and read once, instead of 5 M + the same.
constexpr int CLUSTER_THREADS = 1024;                          // 4 groups of G12::THREADS = 256, one per local block

template <int LOG2N, int LOG2PAD> struct GeomC : GeomL<LOG2N, LOG2PAD> {
    using GL = GeomL<LOG2N, LOG2PAD>;
    static constexpr int CTAS = GL::Q / 4;                     // cluster size: 2, 4 or 8
    static_assert(CTAS >= 2 && CTAS <= 8, "portable cluster sizes");
    static constexpr int SMEM_BYTES = 4 * LARGE_BLOCK * 8;     // 128 KB
    static constexpr int BFLY_PER_CTA = GL::PAD * LARGE_BLOCK / CTAS;
    // columns in pair order: entry 0 = {0, Q/2}, entry e = {e, Q - e}; CTA j holds entries 2j and 2j+1
    static __host__ __device__ constexpr int entry_of_column(int c) { return (c == 0 || c == GL::Q / 2) ? 0 : (c < GL::Q / 2 ? c : GL::Q - c); }
    static __host__ __device__ constexpr int owner_of_column(int c) { return entry_of_column(c) >> 1; }
    static __host__ __device__ constexpr int slot_of_column(int c)
    {
        const int e = entry_of_column(c);
        const int second = e == 0 ? (c == 0 ? 0 : 1) : (c == e ? 0 : 1);
        return ((e & 1) << 1) + second;
    }
    static __host__ __device__ constexpr int column_of_slot(int rank, int slot)
    {
        const int e = 2 * rank + (slot >> 1);
        return e == 0 ? ((slot & 1) ? GL::Q / 2 : 0) : ((slot & 1) ? GL::Q - e : e);
    }
};

// ---- the phases of the cluster kernel as functions of (CTA rank, unit): the kernel below strings them together; the
// host emulation (tests/emu) runs the same code for every rank and thread id in turn ------------------------------------

// phase A, one thread: its share of the first-pass butterflies; every output goes to store(owner CTA, float2 slot in
// that CTA's [4][4096] buffer, value)
template <int LOG2N, int LOG2PAD, int CH, class Store>
__device__ __forceinline__ void cluster_phase_a(const LargeArgs &la, unsigned int rank, unsigned int unit, Store store)
{
    using GC = GeomC<LOG2N, LOG2PAD>;
    using GL = GeomL<LOG2N, LOG2PAD>;
    constexpr int R = GL::R0;
    const KernelArgs &a = la.k;
    const unsigned long long frame0 = la.frame_base + (unsigned long long)unit * (CH == 1 ? 2 : 1);
    const bool second_valid = CH == 1 && frame0 + 1 < a.n_frames;
    const float *pa = a.pcm + frame0 * (unsigned long long)a.hop * CH;
    const float *pb = pa + a.hop;
    // BFLY_PER_CTA / CLUSTER_THREADS butterflies of R inputs per thread = 16 inputs in every geometry: all loads
    // are issued before the first butterfly (one CTA owns the SM here, nothing else hides their latency)
    constexpr int PER_THREAD = GC::BFLY_PER_CTA / CLUSTER_THREADS;
    static_assert(PER_THREAD * CLUSTER_THREADS == GC::BFLY_PER_CTA && PER_THREAD * R == 16, "16 inputs per thread");
    float2 v[PER_THREAD][R];
#pragma unroll
    for (int t = 0; t < PER_THREAD; ++t) {
        const int b = (int)rank * GC::BFLY_PER_CTA + t * CLUSTER_THREADS + (int)threadIdx.x;
        const int q0 = b >> LARGE_LOG2_BLOCK, lo = b & (LARGE_BLOCK - 1);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int n = lo + r * LARGE_BLOCK;
            float2 s;
            if constexpr (CH == 2) {
                s = __ldg(reinterpret_cast<const float2 *>(pa) + n);
            } else {
                s.x = __ldg(pa + n);
                s.y = second_valid ? __ldg(pb + n) : 0.0f;
            }
            // q0 = 0: the plain window (fft.rs:57-63); q0 > 0: window * W_M^(n q0) from one table
            const float2 w = q0 == 0 ? make_float2(__ldg(a.window + n), 0.0f) : __ldg(a.wtw + (size_t)(q0 - 1) * GL::N + n);
            v[t][r] = q0 == 0 ? make_float2(s.x * w.x, s.y * w.x) : cmul(s, w);
        }
    }
#pragma unroll
    for (int t = 0; t < PER_THREAD; ++t) {
        const int b = (int)rank * GC::BFLY_PER_CTA + t * CLUSTER_THREADS + (int)threadIdx.x;
        const int q0 = b >> LARGE_LOG2_BLOCK, lo = b & (LARGE_BLOCK - 1);
        Butterfly<R>::run(v[t]);
        apply_twiddles<R>(v[t], a.twp + lo, LARGE_BLOCK);                     // W_N^(lo q)
        const unsigned int slot_off = (unsigned int)G12::zsw(lo);
#pragma unroll
        for (int q = 0; q < R; ++q) {
            const int c = q * GL::PAD + q0;                                   // column of output d0 = q: bins kk = k Q + c
            const int e = (c == 0 || c == GL::Q / 2) ? 0 : (c < GL::Q / 2 ? c : GL::Q - c);
            const int slot = ((e & 1) << 1) + (e == 0 ? (c == 0 ? 0 : 1) : (c == e ? 0 : 1));
            store((unsigned int)(e >> 1), (unsigned int)slot * (unsigned int)LARGE_BLOCK + slot_off, v[t][q]);
        }
    }
}

// phase B, pass P of the four local 4096-point blocks: thread group g = threadIdx.x / 256 owns block g
template <int P>
__device__ __forceinline__ void cluster_phase_b(float2 *zb, const LargeArgs &la)
{
    KernelArgs ab = la.k;
    ab.twp = la.twp_block;
    const int group = threadIdx.x >> 8, tid = threadIdx.x & 255;
    pass_inplace_t<G12, P>(zb + group * LARGE_BLOCK, ab, tid);
}

// phase C: local L/R split -> global magnitudes by bin kk
template <int LOG2N, int LOG2PAD, int CH>
__device__ __forceinline__ void cluster_phase_c(const float2 *zb, const LargeArgs &la, unsigned int rank, unsigned int unit)
{
    using GC = GeomC<LOG2N, LOG2PAD>;
    using GL = GeomL<LOG2N, LOG2PAD>;
    const KernelArgs &a = la.k;
    const unsigned long long frame0 = la.frame_base + (unsigned long long)unit * (CH == 1 ? 2 : 1);
    const bool second_valid = CH == 1 && frame0 + 1 < a.n_frames;
    float2 *mag_u = la.magbuf + (size_t)unit * (GL::M / 2);
    for (int i = threadIdx.x; i < 4 * (LARGE_BLOCK / 2); i += CLUSTER_THREADS) {
        const int slot = i >> (LARGE_LOG2_BLOCK - 1), k = i & (LARGE_BLOCK / 2 - 1);
        const int c = GC::column_of_slot((int)rank, slot);
        const int kk = k * GL::Q + c;
        if (kk == 0) continue;                                                // DC is dropped (fft.rs:81)
        const bool self = (c == 0 || c == GL::Q / 2);
        const int pslot = self ? slot : (slot ^ 1);
        const int kp = c != 0 ? LARGE_BLOCK - 1 - k : LARGE_BLOCK - k;        // partner index in block Q - c (c = 0: 4096 - k)
        const float2 own = zb[slot * LARGE_BLOCK + G12::zsw(G12::pos_of_bin(k))];
        const float2 par = zb[pslot * LARGE_BLOCK + G12::zsw(G12::pos_of_bin(kp))];
        const float2 m = split_pair(own, par, a.scale);
        mag_u[kk] = m;
        if (a.mags_out != nullptr) {
            float2 *out = reinterpret_cast<float2 *>(a.mags_out) + frame0 * a.k_out + (kk - 1);
            if constexpr (CH == 1) {
                out[0] = make_float2(m.x, m.x);
                if (second_valid) out[a.k_out] = make_float2(m.y, m.y);
            } else {
                out[0] = m;
            }
        }
    }
}

#ifndef SBX_EMU
__device__ __forceinline__ unsigned int cluster_ctarank()
{
    unsigned int r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local_smem_addr` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ unsigned int cluster_map(unsigned int local_smem_addr, unsigned int rank)
{
    unsigned int r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void cluster_store(unsigned int cluster_addr, float2 v)
{
    asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" :: "r"(cluster_addr), "f"(v.x), "f"(v.y) : "memory");
}

// grid = units * CTAS (cluster dimension CTAS), CLUSTER_THREADS threads, dynamic shared memory = GeomC::SMEM_BYTES
template <int LOG2N, int LOG2PAD, int CH>
__global__ void __launch_bounds__(CLUSTER_THREADS, 1) large_cluster_kernel(const LargeArgs la)
{
    using GC = GeomC<LOG2N, LOG2PAD>;
    extern __shared__ __align__(16) unsigned char cluster_smem[];
    float2 *zb = reinterpret_cast<float2 *>(cluster_smem);                    // [4][4096]
    const unsigned int rank = cluster_ctarank();
    const unsigned int unit = blockIdx.x / GC::CTAS;

    // ---- phase A: first-pass butterflies, outputs scattered over the cluster ------------------------------------
    const unsigned int smem_base = (unsigned int)__cvta_generic_to_shared(zb);
    cluster_phase_a<LOG2N, LOG2PAD, CH>(la, rank, unit, [smem_base](unsigned int owner, unsigned int slot, float2 v) {
        cluster_store(cluster_map(smem_base, owner) + slot * 8u, v);
    });
    cluster_sync_all();

    // ---- phase B: four local 4096-point blocks, one thread group each -----------------------------------------
    cluster_phase_b<0>(zb, la);
    __syncthreads();
    cluster_phase_b<1>(zb, la);
    __syncthreads();
    cluster_phase_b<2>(zb, la);
    __syncthreads();

    // ---- phase C: local L/R split -> global magnitudes by bin kk --------------------------------------------------
    cluster_phase_c<LOG2N, LOG2PAD, CH>(zb, la, rank, unit);
}
#endif  // SBX_EMU

#ifndef SBX_EMU
// grid = (units, row chunks): row chunks are independent, so each CTA resamples and colours one chunk of one
// unit; mags by bin kk from global memory (magbuf, or a caller's K x (L,R) array shifted by one)
template <bool PAIR>
__global__ void __launch_bounds__(LARGE_THREADS) large_colour_kernel(const LargeArgs la, const float2 *mags_base, size_t mags_stride)
{
    __shared__ __align__(16) float2 scratch[LARGE_BLOCK];
    const KernelArgs &a = la.k;
    const float2 *mags = mags_base + (size_t)blockIdx.x * mags_stride;
    const unsigned long long frame0 = la.frame_base + (unsigned long long)blockIdx.x * (PAIR ? 2 : 1);
    const bool second_valid = PAIR && frame0 + 1 < a.n_frames;
    const uint4 raw = __ldg(reinterpret_cast<const uint4 *>(a.chunks) + blockIdx.y);
    const RowChunk ck = {raw.x, raw.y, raw.z, raw.w};
    if (ck.samp_end - ck.samp_begin > (uint32_t)LARGE_BLOCK) {
        // one row with more samples than the scratch holds: resample it in parts, one thread carries the sum
        float2 acc = make_float2(0.0f, 0.0f);
        for (uint32_t lo = ck.samp_begin; lo < ck.samp_end; lo += LARGE_BLOCK) {
            const uint32_t hi = lo + LARGE_BLOCK < ck.samp_end ? lo + LARGE_BLOCK : ck.samp_end;
            big_row_part_samples<LARGE_THREADS>(mags, scratch, a, lo, hi);
            __syncthreads();
            big_row_part_sum(scratch, hi - lo, acc);
            __syncthreads();
        }
        big_row_finish<PAIR>(acc, ck.samp_end - ck.samp_begin, ck.row_begin, a, frame0, second_valid);
        return;
    }
    resample_samples<LARGE_THREADS>(mags, scratch, a, ck);
    __syncthreads();
    colour_rows<PAIR, LARGE_THREADS>(scratch, a.thr, a.lut, a, ck, frame0, second_valid);
}
#endif  // SBX_EMU

}  // namespace sbx
