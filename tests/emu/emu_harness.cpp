// emu_harness.cpp -- TEST INFRASTRUCTURE (see emu_shim.h).  Runs one CTA of the fused kernel on
// the host by calling each barrier-delimited phase of stft_kernels.cuh for every thread id in
// turn, with the product's own host tables (sbx_tables.h).
#define SBX_EMU 1
#include "../../spectrogram-rs_b200/csrc/stft_large.cuh"
#include "../../spectrogram-rs_b200/csrc/stft_q2.cuh"
#include "../../spectrogram-rs_b200/csrc/sbx_tables.h"

#include <vector>

namespace {

bool g_emu_staged = false;
bool g_emu_wide = false;                // stereo frames on the WIDE (single-frame latency) geometries
bool g_emu_q2 = false;                  // pad factor 2 on the packed-lane kernel of stft_q2.cuh
std::vector<uint32_t> g_emu_pan_lut;    // stereo schemes: finer pan gradient (sbx_plan_set_pan_gradient), empty = the 256-entry lut
bool g_emu_cluster = false;             // large path: the cluster kernel's phases instead of the four kernels' (M <= 131072)
int g_emu_real = 2;                     // mono batches on the REAL geometry: 0 never, 1 wherever its tables exist, 2 the product's default

template <class G, int P, int END = G::NPASS> struct EmuPasses {
    static void run(float2 *z, const sbx::KernelArgs &a)
    {
        if constexpr (P < END) {
            for (int t = 0; t < G::THREADS; ++t) { threadIdx.x = t; sbx::pass_inplace<G, P>(z, a); }
            EmuPasses<G, P + 1, END>::run(z, a);
        }
    }
};

// one CTA: frames frame0 (.. frame0+1 for mono) -> mags (FRAMES_PER_CTA x K x 2) and rgba
template <class G>
int emu_cta(const sbx::KernelArgs &a, unsigned long long frame0, float *mags_out, uint8_t *rgba)
{
    std::vector<float2> z(G::M);
    const bool second_valid = G::CH == 1 && frame0 + 1 < a.n_frames;
    const float *src = a.pcm + frame0 * (unsigned long long)a.hop * G::CH;
    if (g_emu_staged) {
        // the persistent kernel's path: the unit's sample span sits in a (shared-memory) stage buffer
        const size_t span = ((size_t)G::N + (second_valid ? a.hop : 0)) * G::CH;
        std::vector<float> stage(((size_t)G::N + a.hop) * G::CH, 0.0f);
        std::copy(src, src + span, stage.begin());
        for (int t = 0; t < G::THREADS; ++t) { threadIdx.x = t; sbx::pass_first<G, true>(z.data(), a, stage.data(), second_valid); }
    } else {
        for (int t = 0; t < G::THREADS; ++t) { threadIdx.x = t; sbx::pass_first<G, false>(z.data(), a, src, second_valid); }
    }
    std::vector<sbx::SplitRegs<G>> regs(G::THREADS);
    if constexpr (sbx::LastSplit<G>::ENABLED) {
        EmuPasses<G, 1, G::NPASS - 1>::run(z.data(), a);
        for (int t = 0; t < G::THREADS; ++t) { threadIdx.x = t; sbx::last_pass_keep_own<G>(z.data(), regs[t]); }
        for (int t = 0; t < G::THREADS; ++t) { threadIdx.x = t; sbx::split_with_own<G>(z.data(), a, regs[t]); }
    } else {
        EmuPasses<G, 1>::run(z.data(), a);
        for (int t = 0; t < G::THREADS; ++t) { threadIdx.x = t; sbx::split_read<G>(z.data(), a, regs[t]); }
    }
    for (int t = 0; t < G::THREADS; ++t) { threadIdx.x = t; sbx::split_write<G>(z.data(), regs[t]); }
    if constexpr (G::REAL) {
        const float *zf = reinterpret_cast<const float *>(z.data());
        if (mags_out) {
            float *o = mags_out + (size_t)frame0 * a.k_out * 2;
            for (uint32_t j = 0; j < a.k_out; ++j) { o[2 * j] = zf[j + 1]; o[2 * j + 1] = zf[j + 1]; }
        }
        if (rgba) {
            sbx::KernelArgs b = a;
            b.rgba = rgba;
            float *scratch = reinterpret_cast<float *>(z.data()) + G::M;
            for (uint32_t c = 0; c < a.n_chunks; ++c) {
                const sbx::RowChunk ck = a.chunks[c];
                for (int t = 0; t < G::THREADS; ++t) { threadIdx.x = t; sbx::resample_samples<G::THREADS>(zf, scratch, b, ck); }
                for (int t = 0; t < G::THREADS; ++t) { threadIdx.x = t; sbx::colour_rows_single<G::THREADS>(scratch, b, ck, frame0); }
            }
        }
        return 0;
    }
    if (mags_out) {
        for (uint32_t j = 0; j < a.k_out; ++j) {
            const float2 m = z[j + 1];
            float *o = mags_out + (size_t)frame0 * a.k_out * 2;
            if (G::CH == 1) {
                o[2 * j] = m.x; o[2 * j + 1] = m.x;
                if (second_valid) { o[2 * (a.k_out + j)] = m.y; o[2 * (a.k_out + j) + 1] = m.y; }
            } else { o[2 * j] = m.x; o[2 * j + 1] = m.y; }
        }
    }
    if (rgba) {
        sbx::KernelArgs b = a;
        b.rgba = rgba;
        float2 *scratch = z.data() + G::M / 2;
        for (uint32_t c = 0; c < a.n_chunks; ++c) {
            const sbx::RowChunk ck = a.chunks[c];
            for (int t = 0; t < G::THREADS; ++t) { threadIdx.x = t; sbx::resample_samples<G::THREADS>(z.data(), scratch, b, ck); }
            for (int t = 0; t < G::THREADS; ++t) { threadIdx.x = t; sbx::colour_rows<G::CH == 1, G::THREADS>(scratch, a.thr, a.lut, b, ck, frame0, second_valid); }
        }
    }
    return 0;
}

// one CTA of fused_q2_kernel (stft_q2.cuh): the same barrier-delimited phases, each thread's row kept between them
template <class G, int P, int END> struct EmuQ2Passes {
    static void run(float4 *e, const sbx::KernelArgs &a)
    {
        if constexpr (P < END) {
            for (int t = 0; t < G::THREADS; ++t) { threadIdx.x = t; sbx::q2_pass_inplace<G, P>(e, sbx::q2_pass_twiddles<G, P>(a)); }
            EmuQ2Passes<G, P + 1, END>::run(e, a);
        }
    }
};
template <class G>
int emu_cta_q2(const sbx::KernelArgs &a, unsigned long long frame0, float *mags_out, uint8_t *rgba)
{
    std::vector<float4> e(G::N);
    float2 *z = reinterpret_cast<float2 *>(e.data());
    float *pb = reinterpret_cast<float *>(e.data()) + G::M;                       // upper half (bytes M * 4 ..)
    const bool second_valid = G::CH == 1 && frame0 + 1 < a.n_frames;
    const float *src = a.pcm + frame0 * (unsigned long long)a.hop * G::CH;
    for (int t = 0; t < G::THREADS; ++t) { threadIdx.x = t; sbx::q2_pass_first<G>(e.data(), a, src, second_valid); }
    EmuQ2Passes<G, 1, G::NPASS - 1>::run(e.data(), a);
    struct Row { sbx::cpk v[16]; };
    std::vector<Row> rows(G::THREADS);
    for (int t = 0; t < G::THREADS; ++t) { threadIdx.x = t; sbx::q2_last_load<G>(e.data(), rows[t].v); }
    for (float4 &x : e) x = make_float4(NAN, NAN, NAN, NAN);                      // the buffer is dead once every row is in registers
    for (int t = 0; t < G::THREADS; ++t) { threadIdx.x = t; sbx::q2_partner_write<G>(pb, rows[t].v); }
    if constexpr (G::REAL) {
        float *zf = reinterpret_cast<float *>(e.data());
        for (int t = 0; t < G::THREADS; ++t) { threadIdx.x = t; sbx::q2_split_emit_real<G>(zf, pb, a, rows[t].v); }
        if (mags_out) {
            float *o = mags_out + (size_t)frame0 * a.k_out * 2;
            for (uint32_t j = 0; j < a.k_out; ++j) { o[2 * j] = zf[j + 1]; o[2 * j + 1] = zf[j + 1]; }
        }
        if (rgba) {
            sbx::KernelArgs b = a;
            b.rgba = rgba;
            float *scratch = zf + G::M;
            for (uint32_t c = 0; c < a.n_chunks; ++c) {
                const sbx::RowChunk ck = a.chunks[c];
                for (int t = 0; t < G::THREADS; ++t) { threadIdx.x = t; sbx::resample_samples<G::THREADS>(zf, scratch, b, ck); }
                for (int t = 0; t < G::THREADS; ++t) { threadIdx.x = t; sbx::colour_rows_single<G::THREADS>(scratch, b, ck, frame0); }
            }
        }
        return 0;
    }
    for (int t = 0; t < G::THREADS; ++t) { threadIdx.x = t; sbx::q2_split_emit<G>(z, pb, a, rows[t].v); }
    if (mags_out) {
        for (uint32_t j = 0; j < a.k_out; ++j) {
            const float2 m = z[j + 1];
            float *o = mags_out + (size_t)frame0 * a.k_out * 2;
            if (G::CH == 1) {
                o[2 * j] = m.x; o[2 * j + 1] = m.x;
                if (second_valid) { o[2 * (a.k_out + j)] = m.y; o[2 * (a.k_out + j) + 1] = m.y; }
            } else { o[2 * j] = m.x; o[2 * j + 1] = m.y; }
        }
    }
    if (rgba) {
        sbx::KernelArgs b = a;
        b.rgba = rgba;
        float2 *scratch = z + G::M / 2;
        for (uint32_t c = 0; c < a.n_chunks; ++c) {
            const sbx::RowChunk ck = a.chunks[c];
            for (int t = 0; t < G::THREADS; ++t) { threadIdx.x = t; sbx::pk_resample_samples<G::THREADS, SBX_Q2_RESAMPLE_UNROLL>(z, scratch, b, ck); }
            for (int t = 0; t < G::THREADS; ++t) { threadIdx.x = t; sbx::pk_colour_rows<G::CH == 1, G::THREADS, SBX_Q2_COLOUR_ROWS>(scratch, b, ck, frame0, second_valid); }
        }
    }
    return 0;
}

template <int LOG2N, int LOG2PAD>
int emu_dispatch_ch(int ch, const sbx::KernelArgs &a, unsigned long long f0, float *mags, uint8_t *rgba)
{
    if constexpr (LOG2PAD == 1 && LOG2N >= 9 && LOG2N <= 13)
        if (g_emu_q2) return ch == 1 ? emu_cta_q2<sbx::Q2Geom<LOG2N, 1>>(a, f0, mags, rgba)
                                     : emu_cta_q2<sbx::Q2Geom<LOG2N, 2>>(a, f0, mags, rgba);
    if constexpr (LOG2N + LOG2PAD > 14) return -2;
    else {
        if constexpr (LOG2N + LOG2PAD >= 9 && LOG2N + LOG2PAD <= 12)
            if (ch == 2 && g_emu_wide) return emu_cta<sbx::Geom<LOG2N, LOG2PAD, 2, false, true>>(a, f0, mags, rgba);
        return ch == 1 ? emu_cta<sbx::Geom<LOG2N, LOG2PAD, 1>>(a, f0, mags, rgba)
                       : emu_cta<sbx::Geom<LOG2N, LOG2PAD, 2>>(a, f0, mags, rgba);
    }
}
template <int LOG2N>
int emu_dispatch_pad(int lp, int ch, const sbx::KernelArgs &a, unsigned long long f0, float *mags, uint8_t *rgba)
{
    switch (lp) {
        case 0: return emu_dispatch_ch<LOG2N, 0>(ch, a, f0, mags, rgba);
        case 1: return emu_dispatch_ch<LOG2N, 1>(ch, a, f0, mags, rgba);
        case 2: return emu_dispatch_ch<LOG2N, 2>(ch, a, f0, mags, rgba);
    }
    return -2;
}

template <int LOG2NC, int LOG2PAD>
int emu_real(const sbx::KernelArgs &a, unsigned long long frames, float *mags, uint8_t *rgba)
{
    for (unsigned long long f = 0; f < frames; ++f) {
        if constexpr (LOG2PAD == 1 && LOG2NC >= 9 && LOG2NC <= 13)
            if (g_emu_q2) {
                const int st = emu_cta_q2<sbx::Q2Geom<LOG2NC, 2, true>>(a, f, mags, rgba);
                if (st != 0) return st;
                continue;
            }
        const int st = emu_cta<sbx::Geom<LOG2NC, LOG2PAD, 2, true>>(a, f, mags, rgba);
        if (st != 0) return st;
    }
    return 0;
}

// ---- large path: the four kernels of stft_large.cuh, one block at a time ------------------------------
// the cluster path (large_cluster_kernel): every CTA rank and thread id in turn through the kernel's own phase functions;
// the st.shared::cluster scatter of phase A becomes a store into the owning rank's buffer
template <int LOG2N, int LOG2PAD, int CH>
void emu_cluster_unit(const sbx::LargeArgs &la)
{
    using GL = sbx::GeomL<LOG2N, LOG2PAD>;
    if constexpr (GL::Q / 4 >= 2 && GL::Q / 4 <= 8) {
        using GC = sbx::GeomC<LOG2N, LOG2PAD>;
        std::vector<std::vector<float2>> zb(GC::CTAS, std::vector<float2>(4 * sbx::LARGE_BLOCK));
        for (unsigned rank = 0; rank < (unsigned)GC::CTAS; ++rank)
            for (int t = 0; t < sbx::CLUSTER_THREADS; ++t) {
                threadIdx.x = t;
                sbx::cluster_phase_a<LOG2N, LOG2PAD, CH>(la, rank, 0u, [&zb](unsigned owner, unsigned slot, float2 v) { zb[owner][slot] = v; });
            }
        for (unsigned rank = 0; rank < (unsigned)GC::CTAS; ++rank) {
            for (int t = 0; t < sbx::CLUSTER_THREADS; ++t) { threadIdx.x = t; sbx::cluster_phase_b<0>(zb[rank].data(), la); }
            for (int t = 0; t < sbx::CLUSTER_THREADS; ++t) { threadIdx.x = t; sbx::cluster_phase_b<1>(zb[rank].data(), la); }
            for (int t = 0; t < sbx::CLUSTER_THREADS; ++t) { threadIdx.x = t; sbx::cluster_phase_b<2>(zb[rank].data(), la); }
            for (int t = 0; t < sbx::CLUSTER_THREADS; ++t) { threadIdx.x = t; sbx::cluster_phase_c<LOG2N, LOG2PAD, CH>(zb[rank].data(), la, rank, 0u); }
        }
    }
}

template <int LOG2N, int LOG2PAD, int CH>
int emu_large(sbx::LargeArgs la, float *mags_out, uint8_t *rgba)
{
    using GL = sbx::GeomL<LOG2N, LOG2PAD>;
    const uint32_t per = CH == 1 ? 2 : 1;
    const unsigned long long units = (la.k.n_frames + per - 1) / per;
    std::vector<float2> zbuf(GL::M), magbuf(GL::M / 2);
    la.zbuf = zbuf.data();
    la.magbuf = magbuf.data();
    la.k.mags_out = mags_out;
    la.k.rgba = rgba;
    la.units = 1;
    for (unsigned long long u = 0; u < units; ++u) {
        la.frame_base = u * per;
        const bool cluster = g_emu_cluster && GL::Q / 4 >= 2 && GL::Q / 4 <= 8;
        if (cluster) emu_cluster_unit<LOG2N, LOG2PAD, CH>(la);
        else
        for (unsigned b = 0; b < GL::PAD * (sbx::LARGE_BLOCK / sbx::LARGE_THREADS); ++b)
            for (int t = 0; t < sbx::LARGE_THREADS; ++t) { threadIdx.x = t; sbx::large_first<LOG2N, LOG2PAD, CH>(la, b); }
        for (unsigned b = 0; !cluster && b < GL::M / sbx::LARGE_BLOCK; ++b) {
            std::vector<float2> z(sbx::LARGE_BLOCK);
            float2 *blk = la.zbuf + (size_t)b * sbx::LARGE_BLOCK;
            sbx::KernelArgs a = la.k;
            a.twp = la.twp_block;
            for (int t = 0; t < sbx::LARGE_THREADS; ++t) { threadIdx.x = t; sbx::large_block_load(z.data(), blk, la.twp_block); }
            for (int t = 0; t < sbx::LARGE_THREADS; ++t) { threadIdx.x = t; sbx::pass_inplace<sbx::G12, 1>(z.data(), a); }
            for (int t = 0; t < sbx::LARGE_THREADS; ++t) { threadIdx.x = t; sbx::pass_inplace<sbx::G12, 2>(z.data(), a); }
            for (int t = 0; t < sbx::LARGE_THREADS; ++t) { threadIdx.x = t; sbx::large_block_store(z.data(), blk); }
        }
        for (unsigned b = 0; !cluster && b < (unsigned)GL::Q / 2; ++b) {
            std::vector<float2> own(GL::TK * (GL::Q + 1)), par(GL::TK * (GL::Q + 1));
            for (int t = 0; t < sbx::LARGE_THREADS; ++t) { threadIdx.x = t; sbx::large_split_load<LOG2N, LOG2PAD>(la, b, own.data(), par.data()); }
            for (int t = 0; t < sbx::LARGE_THREADS; ++t) { threadIdx.x = t; sbx::large_split_emit<LOG2N, LOG2PAD, CH>(la, b, own.data(), par.data()); }
        }
        if (rgba) {
            std::vector<float2> scratch(sbx::LARGE_BLOCK);
            const unsigned long long frame0 = la.frame_base;
            const bool second_valid = CH == 1 && frame0 + 1 < la.k.n_frames;
            for (uint32_t c = 0; c < la.k.n_chunks; ++c) {
                const sbx::RowChunk ck = la.k.chunks[c];
                if (ck.samp_end - ck.samp_begin > (uint32_t)sbx::LARGE_BLOCK) {
                    float2 acc = make_float2(0.0f, 0.0f);
                    for (uint32_t lo = ck.samp_begin; lo < ck.samp_end; lo += sbx::LARGE_BLOCK) {
                        const uint32_t hi = std::min(lo + (uint32_t)sbx::LARGE_BLOCK, ck.samp_end);
                        for (int t = 0; t < sbx::LARGE_THREADS; ++t) { threadIdx.x = t; sbx::big_row_part_samples<sbx::LARGE_THREADS>(la.magbuf, scratch.data(), la.k, lo, hi); }
                        threadIdx.x = 0;
                        sbx::big_row_part_sum(scratch.data(), hi - lo, acc);
                    }
                    threadIdx.x = 0;
                    sbx::big_row_finish<CH == 1>(acc, ck.samp_end - ck.samp_begin, ck.row_begin, la.k, frame0, second_valid);
                    continue;
                }
                for (int t = 0; t < sbx::LARGE_THREADS; ++t) { threadIdx.x = t; sbx::resample_samples<sbx::LARGE_THREADS>(la.magbuf, scratch.data(), la.k, ck); }
                for (int t = 0; t < sbx::LARGE_THREADS; ++t) { threadIdx.x = t; sbx::colour_rows<CH == 1, sbx::LARGE_THREADS>(scratch.data(), la.k.thr, la.k.lut, la.k, ck, frame0, second_valid); }
            }
        }
    }
    return 0;
}

template <int LOG2N, int LOG2PAD>
int emu_large_ch(int ch, const sbx::LargeArgs &la, float *mags, uint8_t *rgba)
{
    if constexpr (LOG2N + LOG2PAD <= 14 || LOG2N + LOG2PAD > 18) return -2;
    else return ch == 1 ? emu_large<LOG2N, LOG2PAD, 1>(la, mags, rgba) : emu_large<LOG2N, LOG2PAD, 2>(la, mags, rgba);
}
template <int LOG2N>
int emu_large_pad(int lp, int ch, const sbx::LargeArgs &la, float *mags, uint8_t *rgba)
{
    switch (lp) {
        case 0: return emu_large_ch<LOG2N, 0>(ch, la, mags, rgba);
        case 1: return emu_large_ch<LOG2N, 1>(ch, la, mags, rgba);
        case 2: return emu_large_ch<LOG2N, 2>(ch, la, mags, rgba);
    }
    return -2;
}

// ---- generic (mixed-radix) path: stft_generic.cuh, one CTA ------------------------------------------
template <int CH>
int emu_generic(const sbx::GenericArgs &g0, unsigned long long frames, float *mags_out, uint8_t *rgba)
{
    sbx::GenericArgs g = g0;
    g.k.rgba = rgba;
    const int per = CH == 1 ? 2 : 1;
    for (unsigned long long frame0 = 0; frame0 < frames; frame0 += per) {
        std::vector<float2> z(g.m), mags(g.bl_m / 2);
        const bool second_valid = CH == 1 && frame0 + 1 < frames;
        const bool bluestein = g.bl_bf != nullptr;
        const bool fused_first = !bluestein && SBX_GENERIC_FUSED_FIRST_PASS && (frame0 & 2) == 0;   // both variants get exercised
        for (int t = 0; t < sbx::GENERIC_THREADS; ++t) {
            threadIdx.x = t;
            if (bluestein) sbx::bluestein_load<CH>(z.data(), g, frame0, second_valid);
            else if (fused_first) sbx::generic_first_pass_dispatch<CH>(z.data(), g, frame0, second_valid);
            else sbx::generic_load<CH>(z.data(), g, frame0, second_valid);
        }
        for (uint32_t p = fused_first ? 1u : 0u; p < g.npass; ++p)
            for (int t = 0; t < sbx::GENERIC_THREADS; ++t) { threadIdx.x = t; sbx::generic_pass_dispatch(z.data(), g, p); }
        if (bluestein) {
            for (int t = 0; t < sbx::GENERIC_THREADS; ++t) { threadIdx.x = t; sbx::bluestein_multiply(z.data(), g); }
            for (uint32_t p = g.npass; p-- > 0;)
                for (int t = 0; t < sbx::GENERIC_THREADS; ++t) { threadIdx.x = t; sbx::generic_pass_inverse_dispatch(z.data(), g, p); }
            for (int t = 0; t < sbx::GENERIC_THREADS; ++t) { threadIdx.x = t; sbx::bluestein_split(z.data(), mags.data(), g); }
        } else {
            for (int t = 0; t < sbx::GENERIC_THREADS; ++t) { threadIdx.x = t; sbx::generic_split(z.data(), mags.data(), g); }
        }
        if (mags_out) {
            float *o = mags_out + (size_t)frame0 * g.k.k_out * 2;
            for (uint32_t j = 0; j < g.k.k_out; ++j) {
                const float2 m = mags[j + 1];
                if (CH == 1) {
                    o[2 * j] = m.x; o[2 * j + 1] = m.x;
                    if (second_valid) { o[2 * (g.k.k_out + j)] = m.y; o[2 * (g.k.k_out + j) + 1] = m.y; }
                } else { o[2 * j] = m.x; o[2 * j + 1] = m.y; }
            }
        }
        if (rgba) {
            for (uint32_t c = 0; c < g.k.n_chunks; ++c) {
                const sbx::RowChunk ck = g.k.chunks[c];
                for (int t = 0; t < sbx::GENERIC_THREADS; ++t) { threadIdx.x = t; sbx::resample_samples<sbx::GENERIC_THREADS>(mags.data(), z.data(), g.k, ck); }
                for (int t = 0; t < sbx::GENERIC_THREADS; ++t) { threadIdx.x = t; sbx::colour_rows<CH == 1, sbx::GENERIC_THREADS>(z.data(), g.k.thr, g.k.lut, g.k, ck, frame0, second_valid); }
            }
        }
    }
    return 0;
}

int ilog2(uint32_t v) { int l = 0; while ((1u << l) < v) ++l; return (1u << l) == v ? l : -1; }

}  // namespace

// column -> (owner CTA, slot) assignment of the cluster path (GeomC, stft_large.cuh) for M = 2^log2m: out[3 c] = owner,
// out[3 c + 1] = slot, out[3 c + 2] = column_of_slot(owner, slot); returns Q (columns), 0 for an unsupported length
template <int LOG2N, int LOG2PAD> static int cluster_layout_of(int32_t *out)
{
    using GC = sbx::GeomC<LOG2N, LOG2PAD>;
    for (int c = 0; c < GC::Q; ++c) {
        out[3 * c] = GC::owner_of_column(c);
        out[3 * c + 1] = GC::slot_of_column(c);
        out[3 * c + 2] = GC::column_of_slot(GC::owner_of_column(c), GC::slot_of_column(c));
    }
    return GC::Q;
}

extern "C" {

// Whole batch through the emulated kernel.  pcm: n_samples x channels.  Outputs may be NULL.
// Returns frames processed or a negative error.
__attribute__((visibility("default")))
long long emu_batch(const sbx_config *cfg, const float *pcm, unsigned long long n_samples,
                    float *mags_lr, uint8_t *rgba)
{
    sbx::tables::HostTables t;
    sbx::tables::build(*cfg, t);
    if (n_samples < t.n) return 0;
    const unsigned long long frames = (n_samples - t.n) / cfg->hop + 1;
    sbx::KernelArgs a{};
    a.pcm = pcm;
    if (!t.ok) return -2;
    a.window = t.window.data();
    a.wtw = reinterpret_cast<const float2 *>(t.wtw.data());
    a.twp = reinterpret_cast<const float2 *>(t.twp.data());
    a.wq = reinterpret_cast<const float2 *>(t.wq.data());
    a.rowmap = t.rowmap.data();
    a.rowinfo = t.rowinfo.data();
    a.chunks = t.chunks.data();
    a.n_chunks = (uint32_t)t.chunks.size();
    if (!t.chunks.empty()) a.chunk0 = t.chunks[0];
    a.samp = reinterpret_cast<const uint2 *>(t.samp.data());
    a.taps_clamped = t.taps_clamped ? 1 : 0;
    a.thr = t.thr.data();
    a.lut = t.lut;
    a.ctab = reinterpret_cast<const uint4 *>(t.ctab);
    a.pan_lut = g_emu_pan_lut.empty() ? t.lut : g_emu_pan_lut.data();
    a.pan_levels = g_emu_pan_lut.empty() ? 256u : (uint32_t)g_emu_pan_lut.size();
    a.n_frames = frames;
    a.hop = cfg->hop; a.rows = t.rows; a.k_out = t.k_out;
    a.scale = t.scale; a.db_eps = cfg->db_eps; a.neg_zero = -0.0f; a.step_margin = t.step_margin;
    for (int i = 0; i < 2; ++i) { a.guess_a[i] = t.guess_a[i]; a.guess_b[i] = t.guess_b[i]; }
    a.interp = (int)cfg->interp; a.scheme = (int)cfg->scheme;
    const int ln = ilog2(t.n), lp = ilog2(t.pad), ch = (int)cfg->channels;
    const bool use_real = g_emu_real == 0 ? false : (g_emu_real == 1 ? t.real_ok : t.real_auto);
    if (ch == 1 && use_real && (cfg->hop & 1u) == 0) {
        sbx::KernelArgs r = a;
        r.twp = reinterpret_cast<const float2 *>(t.twp_real.data());
        r.wq = reinterpret_cast<const float2 *>(t.wq_real.data());
        r.wtw4 = reinterpret_cast<const float4 *>(t.wtw4.data());
        r.rtw = reinterpret_cast<const float2 *>(t.rtw.data());
        r.hop = cfg->hop / 2;
        int st = -2;
        switch ((ln + lp) * 4 + lp) {
            case 13 * 4 + 0: st = emu_real<12, 0>(r, frames, mags_lr, rgba); break;
            case 13 * 4 + 1: st = emu_real<11, 1>(r, frames, mags_lr, rgba); break;
            case 13 * 4 + 2: st = emu_real<10, 2>(r, frames, mags_lr, rgba); break;
            case 14 * 4 + 0: st = emu_real<13, 0>(r, frames, mags_lr, rgba); break;
            case 14 * 4 + 1: st = emu_real<12, 1>(r, frames, mags_lr, rgba); break;
            case 14 * 4 + 2: st = emu_real<11, 2>(r, frames, mags_lr, rgba); break;
            case 15 * 4 + 0: st = emu_real<14, 0>(r, frames, mags_lr, rgba); break;
            case 15 * 4 + 1: st = emu_real<13, 1>(r, frames, mags_lr, rgba); break;
            case 15 * 4 + 2: st = emu_real<12, 2>(r, frames, mags_lr, rgba); break;
        }
        return st != 0 ? st : (long long)frames;
    }
    if (t.generic) {
        sbx::GenericArgs g{};
        g.k = a;
        g.twm = reinterpret_cast<const float2 *>(t.twm.data());
        g.pos_tab = t.pos_tab.data();
        sbx::tables::fill_generic_passes(t, g);
        g.twg = reinterpret_cast<const float2 *>(t.twg.data());
        if (t.bluestein) {
            g.bl_a = reinterpret_cast<const float2 *>(t.bl_a.data());
            g.bl_bf = reinterpret_cast<const float2 *>(t.bl_bf.data());
            g.bl_c = reinterpret_cast<const float2 *>(t.bl_c.data());
        }
        const int st = ch == 1 ? emu_generic<1>(g, frames, mags_lr, rgba) : emu_generic<2>(g, frames, mags_lr, rgba);
        return st != 0 ? st : (long long)frames;
    }
    if (t.large) {
        sbx::LargeArgs la{};
        la.k = a;
        la.twp_block = reinterpret_cast<const float2 *>(t.twp_block.data());
        int st;
        switch (ln) {
            case 13: st = emu_large_pad<13>(lp, ch, la, mags_lr, rgba); break;
            case 14: st = emu_large_pad<14>(lp, ch, la, mags_lr, rgba); break;
            case 15: st = emu_large_pad<15>(lp, ch, la, mags_lr, rgba); break;
            case 16: st = emu_large_pad<16>(lp, ch, la, mags_lr, rgba); break;
            default: st = -2;
        }
        return st != 0 ? st : (long long)frames;
    }
    if (ln < 8 || ln > 14 || lp < 0) return -2;
    const int per = ch == 1 ? 2 : 1;
    for (unsigned long long f0 = 0; f0 < frames; f0 += per) {
        int st;
        switch (ln) {
            case 8: st = emu_dispatch_pad<8>(lp, ch, a, f0, mags_lr, rgba); break;
            case 9: st = emu_dispatch_pad<9>(lp, ch, a, f0, mags_lr, rgba); break;
            case 10: st = emu_dispatch_pad<10>(lp, ch, a, f0, mags_lr, rgba); break;
            case 11: st = emu_dispatch_pad<11>(lp, ch, a, f0, mags_lr, rgba); break;
            case 12: st = emu_dispatch_pad<12>(lp, ch, a, f0, mags_lr, rgba); break;
            case 13: st = emu_dispatch_pad<13>(lp, ch, a, f0, mags_lr, rgba); break;
            case 14: st = emu_dispatch_pad<14>(lp, ch, a, f0, mags_lr, rgba); break;
            default: st = -2;
        }
        if (st != 0) return st;
    }
    return (long long)frames;
}

__attribute__((visibility("default")))
void emu_set_staged(int staged) { g_emu_staged = staged != 0; }

__attribute__((visibility("default")))
void emu_set_real(int mode) { g_emu_real = mode; }

__attribute__((visibility("default")))
void emu_set_wide(int wide) { g_emu_wide = wide != 0; }

__attribute__((visibility("default")))
void emu_set_q2(int q2) { g_emu_q2 = q2 != 0; }

__attribute__((visibility("default")))
void emu_set_cluster(int on) { g_emu_cluster = on != 0; }

__attribute__((visibility("default")))
void emu_set_pan_gradient(const uint8_t *rgb, uint32_t entries)
{
    g_emu_pan_lut.clear();
    for (uint32_t i = 0; rgb && i < entries; ++i)
        g_emu_pan_lut.push_back((uint32_t)rgb[3 * i] | ((uint32_t)rgb[3 * i + 1] << 8) | ((uint32_t)rgb[3 * i + 2] << 16) | (255u << 24));
}

// host tables of the product, exposed so CPU tests can compare them with the oracle bit for bit
__attribute__((visibility("default")))
int emu_tables(const sbx_config *cfg, float *window, float *f0, float *step, uint32_t *n, float *period, float *thr512)
{
    sbx::tables::HostTables t;
    sbx::tables::build(*cfg, t);
    if (window) memcpy(window, t.window.data(), t.n * sizeof(float));
    for (uint32_t py = 0; py < t.rows; ++py) {
        if (f0) f0[py] = t.rowmap[py].f0;
        if (step) step[py] = t.rowmap[py].step;
        if (n) n[py] = t.rowmap[py].n;
    }
    if (period) *period = t.period;
    if (thr512) memcpy(thr512, t.thr_raw.data(), 512 * sizeof(float));
    return 0;
}

__attribute__((visibility("default")))
int emu_cluster_layout(int log2m, int32_t *out)
{
    switch (log2m) {
        case 15: return cluster_layout_of<13, 2>(out);
        case 16: return cluster_layout_of<14, 2>(out);
        case 17: return cluster_layout_of<15, 2>(out);
        default: return 0;
    }
}

// the per-sample table, row chunks and row info (n | first slot << 16) of the product, for the CPU test of the
// shared-memory layout rules (conflict-free lane groups); returns the number of table entries, -1 if a buffer is small
__attribute__((visibility("default")))
long long emu_sample_table(const sbx_config *cfg, uint32_t *samp, unsigned long long samp_capacity_entries, uint32_t *chunks,
                           uint32_t max_chunks, uint32_t *n_chunks, uint32_t *rowinfo, uint32_t *bank_group, int *ok)
{
    sbx::tables::HostTables t;
    sbx::tables::build(*cfg, t);
    if (t.samp.size() / 2 > samp_capacity_entries || t.chunks.size() > max_chunks) return -1;
    memcpy(samp, t.samp.data(), t.samp.size() * sizeof(uint32_t));
    for (size_t c = 0; c < t.chunks.size(); ++c) {
        chunks[4 * c] = t.chunks[c].row_begin; chunks[4 * c + 1] = t.chunks[c].row_end;
        chunks[4 * c + 2] = t.chunks[c].samp_begin; chunks[4 * c + 3] = t.chunks[c].samp_end;
    }
    *n_chunks = (uint32_t)t.chunks.size();
    memcpy(rowinfo, t.rowinfo.data(), t.rows * sizeof(uint32_t));
    *bank_group = t.bank_group;
    *ok = t.ok ? 1 : 0;
    return (long long)(t.samp.size() / 2);
}

}  // extern "C"
