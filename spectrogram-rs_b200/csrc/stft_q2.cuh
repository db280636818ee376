// stft_q2.cuh -- the fused frame kernel for pad factor 2, two sub-transforms per thread in packed FP32 lanes.
//
// At P = 2 the zero-padded M-point transform is two N-point sub-transforms (stft_kernels.cuh): q0 = 0 over z[n] w[n]
// (even bins) and q0 = 1 over z[n] w[n] W_M^n (odd bins).  Both run the SAME butterflies with the SAME twiddles, so
// one thread takes butterfly j of both and keeps them in the two lanes of Blackwell's packed FP32 instructions
// (sbx_pk.cuh: FADD2 / FMUL2 / FFMA2) -- every add, multiply, rotation and twiddle of the transform is one
// instruction for the pair, every shared-memory access is 128 bits, PCM / window / twiddle values are loaded once
// instead of twice, and the CTA has N/16 threads with 128 registers instead of N/8 with 64.
//
// Shared-memory element i (16 bytes) = (A[i].re, B[i].re, A[i].im, B[i].im), A = sub-transform 0, B = sub-transform
// 1, i = position in the in-place DIF (digit-reversed on output), XOR-swizzled in its low 3 bits so that the
// strided passes (a quarter warp = 8 consecutive elements = one 128-byte line) and the stride-1 pass (a quarter warp
// = 8 different rows) are both conflict-free.
//
// The last pass is fused with the L/R split as in last_pass_keep_own: thread j takes the row holding bins
// k = j + (N/16) d, d = 0..15, of both sub-transforms -- kk = 2k (A) and 2k + 1 (B).  d < 8 are its own bins; d >= 8 are
// the partners M - kk of other threads' bins: A[N - k] for thread N/16 - j, B[N - 1 - k] for thread N/16 - 1 - j.  They
// are exchanged through four planar float arrays in the upper half of the buffer (32-bit accesses, conflict-free
// on both sides), after which the split itself is lane-uniform again: (L_A, L_B) and (R_A, R_B) come out of the same
// packed instructions, and the magnitudes of bins kk, kk + 1 leave as one 16-byte store in natural bin order -- the
// layout resample_samples / colour_rows of stft_kernels.cuh expect.
#pragma once
#include "sbx_pk.cuh"
#include "stft_kernels.cuh"

#ifndef SBX_Q2_WTW_COMPUTE
#define SBX_Q2_WTW_COMPUTE 1           // pass 0: lane B's factor W_M^n from one table value per thread and constants (else the wtw table)
#endif
#ifndef SBX_Q2_THREAD_BUDGET
#define SBX_Q2_THREAD_BUDGET 768        // 3 CTAs of 256 threads at 80 registers: measured 63.7 vs 58.1 M frames/s (cfg1) against 2 CTAs at 116
#endif

namespace sbx {

// REAL_: the complex sequence is the even/odd packing z[n] = x[2n] + i x[2n+1] of ONE mono frame of 2 N samples (the
// real-sequence geometry of stft_kernels.cuh): CH_ = 2 (sample pairs), one frame per CTA.
template <int LOG2N_, int CH_, bool REAL_ = false> struct Q2Geom {
    using GB = Geom<LOG2N_, 0, CH_>;                       // radix plan / digit reversal of the N-point sub-transform
    static constexpr int LOG2N = LOG2N_, CH = CH_, LOG2PAD = 1, PAD = 2;
    static constexpr bool REAL = REAL_;
    static_assert(!REAL_ || CH_ == 2, "the real-sequence packing reads sample pairs");
    static constexpr int N = 1 << LOG2N, M = 2 * N, LOG2M = LOG2N + 1;
    static constexpr int NPASS = GB::NPASS;
    static constexpr int THREADS = N / 16;                 // one radix-16 butterfly PAIR per thread and pass
    // resident threads per SM the register budget is set for: THREADS * MIN_CTAS = SBX_Q2_THREAD_BUDGET (768 -> 80 registers)
    static constexpr int MIN_CTAS = THREADS >= SBX_Q2_THREAD_BUDGET ? 1 : SBX_Q2_THREAD_BUDGET / THREADS;
    static constexpr int FRAMES_PER_CTA = CH == 1 ? 2 : 1;
    static constexpr int SMEM_BYTES = M * 8;
    static_assert(LOG2N >= 9 && NPASS >= 2, "q2 geometry: N >= 512");
    static __host__ __device__ constexpr int log2r(int p) { return GB::log2r(p); }
    static __host__ __device__ constexpr int log2stride(int p) { return GB::log2stride(p); }
    static __host__ __device__ constexpr int tw_offset(int p) { return GB::tw_offset(p); }
    static __host__ __device__ constexpr int pos_of_k(int k) { return GB::pos_of_bin(k); }      // digit reversal

    // Swizzle key: key bit b = position bit 4 + b, XORed with one of the position bits the three lowest bits of k map to
    // (where those are not bits 4..6 already): the lanes of the stride-1 pass differ exactly in those bits.
    static __host__ __device__ constexpr int kbit_pos(int j)
    {
        const int p = pos_of_k(1 << j);
        int b = 0;
        while ((1 << b) < p) ++b;
        return b;
    }
    static __host__ __device__ constexpr int key_ext_src(int b)
    {
        bool used[3] = {false, false, false};
        int ext[3] = {-1, -1, -1};
        for (int j = 0; j < 3; ++j) {
            const int pb = kbit_pos(j);
            if (pb >= 4 && pb <= 6) used[pb - 4] = true;
        }
        int next = 0;
        for (int j = 0; j < 3; ++j) {
            const int pb = kbit_pos(j);
            if (pb > 6) {
                while (next < 3 && used[next]) ++next;
                if (next < 3) { ext[next] = pb; used[next] = true; }
            }
        }
        return ext[b];
    }
    static __host__ __device__ constexpr int qkey(int i)
    {
        int k = (i >> 4) & 7;
        constexpr int e0 = key_ext_src(0), e1 = key_ext_src(1), e2 = key_ext_src(2);
        if (e0 >= 0) k ^= (i >> (e0 >= 0 ? e0 : 0)) & 1;
        if (e1 >= 0) k ^= ((i >> (e1 >= 0 ? e1 : 0)) & 1) << 1;
        if (e2 >= 0) k ^= ((i >> (e2 >= 0 ? e2 : 0)) & 1) << 2;
        return k;
    }
    // element index -> physical 16-byte slot; XOR-linear: qsw(a | b) = qsw(a) ^ qsw(b) for bit-disjoint a, b
    static __host__ __device__ constexpr int qsw(int i) { return i ^ qkey(i); }
};

// physical slot of element (base + off): `off` a compile-time multiple of 16 whose bits are clear in base, bx = qsw(base)
#define SBX_QAT(G, bx, off) (((bx) ^ G::qkey(off)) + (off))

// W_32^k = exp(-2 pi i k / 32), k = 0..15, as compile-time constants
__host__ __device__ constexpr float2 w32_pow_c(int k)
{
    return k < 8 ? float2{w32_cos(k), -w32_sin(k)} : float2{-w32_sin(k - 8), -w32_cos(k - 8)};      // k >= 8: times -i
}
__device__ __forceinline__ float2 w32_pow(int k) { return make_float2(w32_pow_c(k).x, w32_pow_c(k).y); }

// ---- pass 0: PCM -> window (A) / window * W_M^n (B) -> radix butterfly -> twiddle -> shared memory ----------------
// SHIFT > 0 (mono pairs): hop == SHIFT * S, frame B's input r is frame A's input r + SHIFT (R + SHIFT loads, not 2 R)
template <class G, int SHIFT>
__device__ __forceinline__ void q2_first_inputs(cpk (&v)[1 << G::log2r(0)], const KernelArgs &a, const float *pa_lo,
                                                const float *pb_lo, bool second_valid, int lo)
{
    constexpr int R = 1 << G::log2r(0), S = 1 << G::log2stride(0);
    const float *win_lo = a.window + lo;
#if SBX_Q2_WTW_COMPUTE
    const float2 wl = ldg_tab(a.wq + lo);                        // W_M^lo
#else
    const float2 wtw_lo_unused = make_float2(0.0f, 0.0f); (void)wtw_lo_unused;
    const float2 *wtw_lo = a.wtw + lo;                           // window[n] * W_M^n: the q0 = 1 factor
#endif
    float xs[SHIFT > 0 ? R + SHIFT : 1];
    if constexpr (SHIFT > 0) {
#pragma unroll
        for (int r = 0; r < R + SHIFT; ++r)
            xs[r] = (r < R || second_valid) ? ldg_pcm(pa_lo + r * S) : 0.0f;
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        float2 s;
        if constexpr (SHIFT > 0) {
            s.x = xs[r];
            s.y = second_valid ? xs[r + SHIFT] : 0.0f;
        } else if constexpr (G::CH == 2) {
            s = ldg_pcm2(reinterpret_cast<const float2 *>(pa_lo) + r * S);
        } else {
            s.x = ldg_pcm(pa_lo + r * S);
            s.y = second_valid ? ldg_pcm(pb_lo + r * S) : 0.0f;
        }
#if SBX_Q2_WTW_COMPUTE
        // lane A = the windowed sample, lane B = A * W_M^n with W_M^n = W_M^lo * W_(2R)^r: one table value per thread
        // (wl) times compile-time constants instead of a 32-48 KB window * W table streamed through L1 by every CTA
        float2 A;
        if constexpr (G::REAL) {
            const float2 wn = ldg_stream(reinterpret_cast<const float2 *>(a.window) + lo + r * S);
            A = make_float2(s.x * wn.x, s.y * wn.y);             // each sample with its own window value
        } else {
            const float wn = ldg_stream(win_lo + r * S);         // fft.rs:57-63 (complex * f32)
            A = make_float2(s.x * wn, s.y * wn);
        }
        const float2 wr = r == 0 ? wl : cmul(wl, w32_pow(r * (16 / R)));
        v[r].re = pk_make(A.x, A.x * wr.x - A.y * wr.y);
        v[r].im = pk_make(A.y, A.x * wr.y + A.y * wr.x);
#else
        if constexpr (G::REAL) {
            // s = (x[2n], x[2n+1]); A: each sample with its own window value; B: times W_(M)^n as well (wtw4, sbx_tables.h)
            const float2 wn = ldg_stream(reinterpret_cast<const float2 *>(a.window) + lo + r * S);
            const float4 f = ldg_stream(a.wtw4 + lo + r * S);
            v[r].re = pk_make(s.x * wn.x, s.x * f.x - s.y * f.y);
            v[r].im = pk_make(s.y * wn.y, s.x * f.z + s.y * f.w);
        } else {
            const float wn = ldg_stream(win_lo + r * S);         // fft.rs:57-63 (complex * f32)
            const float2 wc = ldg_stream(wtw_lo + r * S);
            v[r].re = pk_make(s.x * wn, s.x * wc.x - s.y * wc.y);
            v[r].im = pk_make(s.y * wn, s.x * wc.y + s.y * wc.x);
        }
#endif
    }
}

template <class G>
__device__ __forceinline__ void q2_pass_first(float4 *e, const KernelArgs &a, const float *pa, bool second_valid)
{
    constexpr int LR = G::log2r(0), R = 1 << LR, LS = G::log2stride(0), S = 1 << LS;
    for (int lo = threadIdx.x; lo < S; lo += G::THREADS) {
        const float *pa_lo = pa + (size_t)lo * G::CH;
        const float *pb_lo = pa_lo + a.hop;
        cpk v[R];
        if constexpr (G::CH == 1 && R >= 8) {
            const uint32_t hop = a.hop;                          // the common overlaps: hop = S or 2 S (cfg1, cfg4)
            if (hop == 2u * S) q2_first_inputs<G, 2>(v, a, pa_lo, pb_lo, second_valid, lo);
            else if (hop == 1u * S) q2_first_inputs<G, 1>(v, a, pa_lo, pb_lo, second_valid, lo);
            else q2_first_inputs<G, 0>(v, a, pa_lo, pb_lo, second_valid, lo);
        } else {
            q2_first_inputs<G, 0>(v, a, pa_lo, pb_lo, second_valid, lo);
        }
        PkButterfly<R>::run(v);
        const cpk w1 = cpk_dup(ldg_tab(a.twp + lo));             // W_N^lo
        const cpk w4 = R >= 8 ? cpk_dup(ldg_tab(a.twp + 2 * S + lo)) : w1;
        pk_apply_twiddles<R>(v, w1, w4);
        const int bx = G::qsw(lo);
#pragma unroll
        for (int q = 0; q < R; ++q) cpk_sts(e + SBX_QAT(G, bx, q * S), v[q]);
    }
}

// ---- passes 1 .. NPASS-2: in-place radix-16 at stride S = 2^LS, one butterfly pair per thread ---------------------
// The pass's two twiddle values do not depend on the data: q2_pass_twiddles is called BEFORE the barrier that precedes
// the pass, so that their latency (an L2 round trip: three resident CTAs leave little L1) overlaps the barrier wait.
struct Q2Twiddles { float2 w1, w4; };
template <class G, int P>
__device__ __forceinline__ Q2Twiddles q2_pass_twiddles(const KernelArgs &a)
{
    constexpr int LS = G::log2stride(P), S = 1 << LS;
    const float2 *tw = a.twp + G::tw_offset(P) + (threadIdx.x & (S - 1));      // W_(16 S)^(lo q)
    Q2Twiddles t;
    t.w1 = ldg_tab(tw);
    t.w4 = ldg_tab(tw + 2 * S);
    return t;
}
template <class G, int P>
__device__ __forceinline__ void q2_pass_inplace(float4 *e, const Q2Twiddles &t)
{
    constexpr int LS = G::log2stride(P), S = 1 << LS;
    static_assert(S >= 16, "the stride-1 pass is q2_last_load");
    const int u = threadIdx.x;
    const int lo = u & (S - 1), hi = u >> LS;
    const int bx = G::qsw((hi << (LS + 4)) + lo);
    cpk v[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) v[r] = cpk_lds(e + SBX_QAT(G, bx, r * S));
    PkButterfly<16>::run(v);
    pk_apply_twiddles<16>(v, cpk_dup(t.w1), cpk_dup(t.w4));
#pragma unroll
    for (int q = 0; q < 16; ++q) cpk_sts(e + SBX_QAT(G, bx, q * S), v[q]);
}

// pass 0's barrier, then passes P .. END-1, each followed by its barrier
template <class G, int P, int END> struct Q2RunPasses {
    static __device__ __forceinline__ void run(float4 *e, const KernelArgs &a)
    {
        if constexpr (P < END) {
            const Q2Twiddles t = q2_pass_twiddles<G, P>(a);
            __syncthreads();
            q2_pass_inplace<G, P>(e, t);
            Q2RunPasses<G, P + 1, END>::run(e, a);
        } else {
            __syncthreads();
        }
    }
};

// ---- last pass (stride 1) fused with the L/R split -------------------------------------------------------------------
// phase 1: thread j loads the row of bins k = j + THREADS d and runs the butterfly: v[d] = (A[k], B[k])
template <class G>
__device__ __forceinline__ void q2_last_load(const float4 *e, cpk (&v)[16])
{
    const int row = G::pos_of_k(threadIdx.x) >> 4;
    const float4 *rowp = e + (row << 4);
    const int key = G::qkey(row << 4);
#pragma unroll
    for (int r = 0; r < 16; ++r) v[r] = cpk_lds(rowp + (r ^ key));
    PkButterfly<16>::run(v);
}

// phase 2 (every row is in registers): bins d >= 8 go to the threads whose partners they are.  pb = four planes of
// 8 x THREADS floats: A.re, A.im, B.re, B.im of reader jr's partner for its own bin d at [d * THREADS + jr].
template <class G>
__device__ __forceinline__ void q2_partner_write(float *pb, const cpk (&v)[16])
{
    constexpr int NT = G::THREADS;
    const int j = threadIdx.x;
    // A[k], k = j + NT e, is the partner of A[N - k] = bin (NT - j) + NT (15 - e); for j = 0: bin 0 + NT (16 - e), and
    // e = 8 is the Nyquist bin kk = M/2 (dropped, fft.rs:33)
    const int ja = ((NT - j) & (NT - 1)) + (j == 0 ? NT : 0);    // j = 0: one digit up
    // B[k] is the partner of B[N - 1 - k] = bin (NT - 1 - j) + NT (15 - e)
    const int jb = NT - 1 - j;
#pragma unroll
    for (int e = 8; e < 16; ++e) {
        const int d = 15 - e;
        if (e != 8 || j != 0) {
            pb[d * NT + ja] = pk_lo(v[e].re);
            pb[8 * NT + d * NT + ja] = pk_lo(v[e].im);
        }
        pb[16 * NT + d * NT + jb] = pk_hi(v[e].re);
        pb[24 * NT + d * NT + jb] = pk_hi(v[e].im);
    }
}

// phase 3: own bins d < 8 with their partners -> (L, R) * 2/N for bins kk = 2k (A) and 2k + 1 (B), written over the
// lower half of the buffer in natural bin order: mags[kk] = (L, R), two bins per 16-byte store
//   L = |a + conj b| / 2 * (2/N),  R = |a - conj b| / 2 * (2/N),  a = Z[kk], b = Z[M - kk]       (fft.rs:81-95)
template <class G>
__device__ __forceinline__ void q2_split_emit(float2 *mags, const float *pb, const KernelArgs &a, const cpk (&v)[16])
{
    constexpr int NT = G::THREADS;
    const int j = threadIdx.x;
    const float hs = 0.5f * a.scale;                             // exact: a power-of-two scaling of 2/N
    float4 *out = reinterpret_cast<float4 *>(mags);
#pragma unroll
    for (int d = 0; d < 8; ++d) {
        const pk yre = pk_make(pb[d * NT + j], pb[16 * NT + d * NT + j]);
        const pk yim = pk_make(pb[8 * NT + d * NT + j], pb[24 * NT + d * NT + j]);
        const pk lx = pk_add(v[d].re, yre), ly = pk_sub(v[d].im, yim);     // a + conj b
        const pk rx = pk_sub(v[d].re, yre), ry = pk_add(v[d].im, yim);     // a - conj b
        const pk l2 = pk_fma(lx, lx, pk_mul(ly, ly)), r2 = pk_fma(rx, rx, pk_mul(ry, ry));
        // kk = 0 (j = 0, d = 0, lane A) is DC: its partner slot was never written and its magnitude is never read
        out[j + NT * d] = make_float4(fast_sqrt(pk_lo(l2)) * hs, fast_sqrt(pk_lo(r2)) * hs,
                                      fast_sqrt(pk_hi(l2)) * hs, fast_sqrt(pk_hi(r2)) * hs);
    }
}

// REAL: a = Z[kk], b = Z[M - kk] of the packed sequence, w = W_(2M)^kk.  With E, O the spectra of the even and odd
// samples, a + conj b = 2 E, a - conj b = 2i O, X[kk] = E + w O and X[M - kk] = conj(E - w O) (split_pair_real).  One
// magnitude per bin of the real spectrum, natural order, over the lower half of the buffer: zf[kk] = |X[kk]| * 2/N.
// Thread 0 also owns the self-paired bin M/2 (its row's element 8, lane A): |X[M/2]| = |Z[M/2]|.
template <class G>
__device__ __forceinline__ void q2_split_emit_real(float *zf, const float *pb, const KernelArgs &a, const cpk (&v)[16])
{
    constexpr int NT = G::THREADS;
    const int j = threadIdx.x;
    const float hs = 0.5f * a.scale;
    // W_(2M)^kk for kk = 2j (lane A), 2j + 1 (lane B); bins of digit d: times W_32^d
    const float4 wt = ldg_tab(reinterpret_cast<const float4 *>(a.rtw) + j);
    const cpk w0 = cpk_make(pk_make(wt.x, wt.z), pk_make(wt.y, wt.w));
#pragma unroll
    for (int d = 0; d < 8; ++d) {
        const pk yre = pk_make(pb[d * NT + j], pb[16 * NT + d * NT + j]);
        const pk yim = pk_make(pb[8 * NT + d * NT + j], pb[24 * NT + d * NT + j]);
        const cpk w = d == 0 ? w0 : pk_rot(w0, pk_dup(w32_cos(d)), pk_dup(w32_sin(d)));     // w0 * (cos - i sin)
        const pk sx = pk_add(v[d].re, yre), sy = pk_sub(v[d].im, yim);     // a + conj b
        const pk dx = pk_sub(v[d].re, yre), dy = pk_add(v[d].im, yim);     // a - conj b
        const pk tx = pk_fma(w.re, dy, pk_mul(w.im, dx));                  // w * (-i d)
        const pk ty = pk_fma(w.im, dy, pk_neg(pk_mul(w.re, dx)));
        const pk ox = pk_add(sx, tx), oy = pk_add(sy, ty), px = pk_sub(sx, tx), py = pk_sub(sy, ty);
        const pk o2 = pk_fma(ox, ox, pk_mul(oy, oy)), p2 = pk_fma(px, px, pk_mul(py, py));
        const int kk = 2 * (j + NT * d);
        // X[kk], X[kk + 1]; kk = 0 is DC (dropped, never read)
        *reinterpret_cast<float2 *>(zf + kk) = make_float2(fast_sqrt(pk_lo(o2)) * hs, fast_sqrt(pk_hi(o2)) * hs);
        // X[M - kk], X[M - kk - 1]; M - 0 is the Nyquist bin (dropped) and lies outside the magnitudes
        if (d != 0 || j != 0) zf[G::M - kk] = fast_sqrt(pk_lo(p2)) * hs;
        zf[G::M - kk - 1] = fast_sqrt(pk_hi(p2)) * hs;
    }
    if (j == 0) {
        const float re = pk_lo(v[8].re), im = pk_lo(v[8].im);
        zf[G::M / 2] = fast_sqrt(re * re + im * im) * a.scale;
    }
}

#ifndef SBX_Q2_RESAMPLE_UNROLL
#define SBX_Q2_RESAMPLE_UNROLL 8       // sample-table loads in flight per thread (128 registers to spend)
#endif
#ifndef SBX_Q2_COLOUR_ROWS
#define SBX_Q2_COLOUR_ROWS 4           // rows in flight per thread in the colour phase
#endif

#ifndef SBX_Q2_SINGLE_UNROLL
#define SBX_Q2_SINGLE_UNROLL 8         // real-sequence geometry (one float per bin): sample-table loads in flight per thread
#endif
#ifndef SBX_Q2_SINGLE_ROWS
#define SBX_Q2_SINGLE_ROWS 2           // ... and rows in flight per thread in its colour phase (B200, cfg3: (4,2) 34.38, (8,2) 34.63, (8,4) 34.14 M frames/s: 4 rows spill)
#endif

#ifndef SBX_EMU
// -DSBX_PHASE_CLOCK (tools/phase_split.py builds that variant; never the shipped library): thread 0 of every CTA reads
// the SM clock at the phase boundaries and adds the differences to g_phase_cycles -- the per-phase share of a CTA's
// residency, profiles/phase_split_*.json.  [0] pass 0 (PCM + window loads, first radix pass), [1] shared-memory
// passes, [2] last pass + partner exchange + L/R split, [3] resample (lane per sample), [4] row sums + colour + RGBA
// stores, [5] magnitude output, [7] CTAs counted.
#ifdef SBX_PHASE_CLOCK
__device__ unsigned long long g_phase_cycles[8];
#define SBX_PHASE_MARK(i)                                                                   \
    do {                                                                                    \
        if (threadIdx.x == 0) {                                                             \
            const long long now__ = clock64();                                              \
            atomicAdd(&g_phase_cycles[i], (unsigned long long)(now__ - phase_t__));         \
            phase_t__ = now__;                                                              \
        }                                                                                   \
    } while (0)
#define SBX_PHASE_BEGIN() long long phase_t__ = clock64()
#else
#define SBX_PHASE_MARK(i) do { } while (0)
#define SBX_PHASE_BEGIN() do { } while (0)
#endif

template <class G>
__global__ void __launch_bounds__(G::THREADS, G::MIN_CTAS) fused_q2_kernel(const KernelArgs a)
{
    SBX_PHASE_BEGIN();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 *e = reinterpret_cast<float4 *>(smem_raw);
    float2 *z = reinterpret_cast<float2 *>(smem_raw);
    float *pb = reinterpret_cast<float *>(smem_raw + (size_t)G::M * 4);       // upper half: partner planes, then scratch

    unsigned long long unit = blockIdx.x, seq = 0;
    if (a.units_per_seq != 0) {                                   // planar sequences: CTA -> (sequence, unit within it)
        seq = unit / a.units_per_seq;
        unit -= seq * a.units_per_seq;
    }
    const unsigned long long frame0 = unit * G::FRAMES_PER_CTA;
    const bool second_valid = G::CH == 1 && frame0 + 1 < a.n_frames;

    q2_pass_first<G>(e, a, a.pcm + seq * a.seq_stride + frame0 * (unsigned long long)a.hop * G::CH, second_valid);
    SBX_PHASE_MARK(0);
    Q2RunPasses<G, 1, G::NPASS - 1>::run(e, a);              // (each pass is preceded by its barrier; one more at the end)
    SBX_PHASE_MARK(1);
    cpk v[16];
    q2_last_load<G>(e, v);
    __syncthreads();
    q2_partner_write<G>(pb, v);
    __syncthreads();
    if constexpr (G::REAL) q2_split_emit_real<G>(reinterpret_cast<float *>(smem_raw), pb, a, v);
    else q2_split_emit<G>(z, pb, a, v);
    __syncthreads();
    SBX_PHASE_MARK(2);

    const unsigned long long out_frame0 = seq * a.n_frames + frame0;
    if constexpr (G::REAL) {
        const float *zf = reinterpret_cast<const float *>(smem_raw);
        if (a.mags_out != nullptr) {
            float2 *out = reinterpret_cast<float2 *>(a.mags_out) + out_frame0 * a.k_out;
            for (uint32_t j = threadIdx.x; j < a.k_out; j += G::THREADS) out[j] = make_float2(zf[j + 1], zf[j + 1]);
        }
        if (a.rgba != nullptr) resample_colour_single<G::LOG2M, G::THREADS, SBX_Q2_SINGLE_UNROLL, SBX_Q2_SINGLE_ROWS>(z, a, out_frame0);
    } else {
    if (a.mags_out != nullptr) {
        float2 *out = reinterpret_cast<float2 *>(a.mags_out) + out_frame0 * a.k_out;
        for (uint32_t j = threadIdx.x; j < a.k_out; j += G::THREADS) {
            const float2 m = z[j + 1];
            if constexpr (G::CH == 1) {
                out[j] = make_float2(m.x, m.x);
                if (second_valid) out[a.k_out + j] = make_float2(m.y, m.y);
            } else {
                out[j] = m;
            }
        }
    }
    SBX_PHASE_MARK(5);
#ifdef SBX_PHASE_CLOCK
    if (a.rgba != nullptr) {                                   // pk_resample_colour, with a mark between its two phases
        float2 *scratch = z + (1 << (G::LOG2M - 1));
        for (uint32_t c = 0; c < a.n_chunks; ++c) {
            RowChunk ck = a.chunk0;
            if (c != 0) {
                const uint4 raw = __ldg(reinterpret_cast<const uint4 *>(a.chunks) + c);
                ck = RowChunk{raw.x, raw.y, raw.z, raw.w};
                __syncthreads();
            }
            pk_resample_samples<G::THREADS, SBX_Q2_RESAMPLE_UNROLL>(z, scratch, a, ck);
            __syncthreads();
            SBX_PHASE_MARK(3);
            pk_colour_rows<G::CH == 1, G::THREADS, SBX_Q2_COLOUR_ROWS>(scratch, a, ck, out_frame0, second_valid);
            SBX_PHASE_MARK(4);
        }
    }
    if (threadIdx.x == 0) atomicAdd(&g_phase_cycles[7], 1ull);
#else
    if (a.rgba != nullptr) pk_resample_colour<G::LOG2M, G::CH == 1, G::THREADS, SBX_Q2_RESAMPLE_UNROLL, SBX_Q2_COLOUR_ROWS>(z, a, out_frame0, second_valid);
#endif
    }  // !REAL
    if (a.done_flag != nullptr) {
        __syncthreads();
        if (threadIdx.x == 0) {
            const unsigned int v2 = *a.done_count + 1u;
            *a.done_count = v2;
            __threadfence_system();
            *reinterpret_cast<volatile unsigned int *>(a.done_flag) = v2;
        }
    }
}
#endif  // SBX_EMU

}  // namespace sbx
