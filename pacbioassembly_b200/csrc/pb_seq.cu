// pb_seq.cu -- L0: sequence representation on the device (replaces src/dna_seq.h).
//
// Ingest kernel: text (or the reference's packed .bin records) -> padded line in three views
// (2-bit packed bytes, hi/lo bit planes).  One warp produces one 32-base word of the line:
// coalesced 1 B/base reads, ballot-built plane words, shuffle-combined packed words.
#include <algorithm>

#include <chrono>

#include "pb_internal.cuh"

// C2I, dna_seq.h:21 : A->0 C->1 G->2 anything else->3
__device__ __forceinline__ uint32_t c2i(uint32_t ch) { return ch == 'A' ? 0u : ch == 'C' ? 1u : ch == 'G' ? 2u : 3u; }

struct IngestSrc {
    const uint8_t *text;    // device blob
    const int64_t *toff;    // [n] offset of element 0 (text mode) or of the record's first body byte (packed mode)
    const int32_t *tstride; // [n] +1/-1, or NULL
    int packed;             // PB_SRC_TEXT / PB_SRC_PACKED / PB_SRC_REVLINE
};

#define INGEST_RUN 64 // consecutive 32-base words handled by one warp: one owner search, then a forward walk

__global__ void __launch_bounds__(256)
ingest_kernel(IngestSrc src, const int64_t *__restrict__ base, const int32_t *__restrict__ len, int64_t n,
              int64_t nwords, uint32_t *__restrict__ hi, uint32_t *__restrict__ lo, uint32_t *__restrict__ packed,
              uint32_t *__restrict__ irr, uint32_t *__restrict__ flags, unsigned long long *__restrict__ nirr)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t nruns = (nwords + INGEST_RUN - 1) / INGEST_RUN;
    for (int64_t run = warp0; run < nruns; run += nwarps) {
        const int64_t w_begin = run * INGEST_RUN, w_end = min(nwords, w_begin + INGEST_RUN);
        // sequence that owns the first word: largest i with base[i] <= g0 (bases are multiples of 32, base[n] = end)
        int64_t a = 0, b = n;
        while (b - a > 1) {
            int64_t m = (a + b) >> 1;
            if (__ldg(base + m) <= w_begin * 32) a = m; else b = m;
        }
        int64_t a_base = n > 0 ? __ldg(base + a) : 0, a_next = n > 0 ? __ldg(base + a + 1) : 0;
        int64_t a_len = n > 0 ? __ldg(len + a) : 0, a_off = n > 0 ? __ldg(src.toff + a) : 0;
        int64_t a_st = (n > 0 && src.tstride) ? (int64_t)__ldg(src.tstride + a) : 1;
        uint32_t irr_acc = 0u;
        for (int64_t w = w_begin; w < w_end; ++w) {
            const int64_t g0 = w * 32;
            while (n > 0 && a + 1 < n && g0 >= a_next) { // walk to the next sequence (warp-uniform)
                if (irr_acc && lane == 0) atomicOr(&flags[a], PB_FLAG_IRREGULAR);
                irr_acc = 0u;
                ++a;
                a_base = a_next; a_next = __ldg(base + a + 1);
                a_len = __ldg(len + a); a_off = __ldg(src.toff + a);
                a_st = src.tstride ? (int64_t)__ldg(src.tstride + a) : 1;
            }
            const int64_t rel = g0 + lane - a_base;
            const bool valid = n > 0 && rel < a_len;
            uint32_t code = 3u;
            bool irregular = false;
            if (valid) {
                if (src.packed == PB_SRC_PACKED) {
                    uint32_t byte = src.text[a_off + (rel >> 2)];
                    code = (byte >> (6 - 2 * (rel & 3))) & 3u;
                } else if (src.packed == PB_SRC_REVLINE) {
                    const int64_t gs = a_off + (a_len - 1 - rel); // same sequence of the source line, last base first
                    uint32_t byte = src.text[gs >> 2];
                    code = (byte >> (6 - 2 * (gs & 3))) & 3u;
                } else {
                    uint32_t ch = src.text[a_off + rel * a_st];
                    code = c2i(ch);
                    irregular = !(ch == 'A' || ch == 'C' || ch == 'G' || ch == 'T');
                }
            }
            const uint32_t whi = __ballot_sync(0xffffffffu, code & 2u);
            const uint32_t wlo = __ballot_sync(0xffffffffu, code & 1u);
            const uint32_t wirr = __ballot_sync(0xffffffffu, irregular);
            irr_acc |= wirr;
            // packed: byte (lane>>2) of the 8 output bytes; first base of a byte in bits 7:6; bytes little-endian in u32
            uint32_t v = code << (6 - 2 * (lane & 3));
            v <<= 8 * ((lane >> 2) & 3);
            v |= __shfl_xor_sync(0xffffffffu, v, 1);
            v |= __shfl_xor_sync(0xffffffffu, v, 2);
            v |= __shfl_xor_sync(0xffffffffu, v, 4);
            v |= __shfl_xor_sync(0xffffffffu, v, 8);
            if (lane == 0) {
                hi[w] = whi;
                lo[w] = wlo;
                irr[w] = wirr;
                packed[2 * w] = v;
                if (wirr) atomicAdd(nirr, (unsigned long long)__popc(wirr));
            }
            if (lane == 16) packed[2 * w + 1] = v;
        }
        if (irr_acc && lane == 0 && n > 0) atomicOr(&flags[a], PB_FLAG_IRREGULAR);
    }
}

// Text input with stride +1 (the read batches of the locate path): 16 bases per thread.  A thread's granule is 16 consecutive
// bases of the padded line -- sequences start at multiples of 32, so a granule never straddles two of them -- loaded as five
// aligned 32-bit words of the text and re-aligned with funnel shifts; C2I (dna_seq.h:21: A, C, G -> 0, 1, 2, anything else 3) and
// the "not one of ACGT" test are done four bytes at a time with SIMD compares; the 2-bit codes are gathered into the plane bits
// and the packed byte by one multiplication each (the fields of the products do not overlap, so no carries).  A warp covers 512
// bases per step and INGEST_STEPS steps: one owner search per warp, then every lane walks forward on its own.
#define INGEST_STEPS 16
__global__ void __launch_bounds__(256)
ingest_text_kernel(const uint8_t *__restrict__ text, int64_t text_bytes, const int64_t *__restrict__ toff,
                   const int64_t *__restrict__ base, const int32_t *__restrict__ len, int64_t n, int64_t nwords,
                   uint32_t *__restrict__ hi, uint32_t *__restrict__ lo, uint32_t *__restrict__ packed, uint32_t *__restrict__ irr,
                   uint32_t *__restrict__ flags, unsigned long long *__restrict__ nirr)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t g_first = warp * (512 * INGEST_STEPS), g_line = nwords * 32;
    if (g_first >= g_line || n <= 0) return;
    // sequence that owns the warp's first base: largest a with base[a] <= g_first
    int64_t a = 0, b = n;
    while (b - a > 1) {
        const int64_t m = (a + b) >> 1;
        if (__ldg(base + m) <= g_first) a = m; else b = m;
    }
    int64_t a_base = __ldg(base + a), a_next = __ldg(base + a + 1), a_len = __ldg(len + a), a_off = __ldg(toff + a);
    unsigned long long bad = 0ull;
    for (int step = 0; step < INGEST_STEPS; ++step) {
        const int64_t g = g_first + 512 * step + 16 * lane; // first base of this lane's granule
        if (g_first + 512 * step >= g_line) break;
        while (a + 1 < n && g >= a_next) {
            ++a;
            a_base = a_next; a_next = __ldg(base + a + 1);
            a_len = __ldg(len + a); a_off = __ldg(toff + a);
        }
        const int64_t rel = g - a_base;
        const int nvalid = (int)max((int64_t)0, min((int64_t)16, a_len - rel)); // bases of the granule inside the sequence
        uint32_t x[4] = {0x54545454u, 0x54545454u, 0x54545454u, 0x54545454u};   // 'T': code 3 and regular, like the padding
        if (nvalid > 0 && g < g_line) {
            const int64_t p = a_off + rel;
            if ((p & ~(int64_t)3) + 20 <= text_bytes) { // five aligned words around [p, p + 16)
                const uint32_t *w = reinterpret_cast<const uint32_t *>(text + (p & ~(int64_t)3));
                const uint32_t w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2), w3 = __ldg(w + 3), w4 = __ldg(w + 4);
                const unsigned sh = 8u * (unsigned)(p & 3);
                x[0] = __funnelshift_r(w0, w1, sh); x[1] = __funnelshift_r(w1, w2, sh);
                x[2] = __funnelshift_r(w2, w3, sh); x[3] = __funnelshift_r(w3, w4, sh);
            } else { // the last bytes of the blob: one by one
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    uint32_t v = 0u;
#pragma unroll
                    for (int q = 0; q < 4; ++q) v |= (uint32_t)(4 * k + q < nvalid ? text[p + 4 * k + q] : (uint8_t)'T') << (8 * q);
                    x[k] = v;
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) { // bytes past the sequence's end read 'T'
                const int vk = nvalid - 4 * k;
                const uint32_t keep = vk >= 4 ? 0xffffffffu : (vk <= 0 ? 0u : (1u << (8 * vk)) - 1u);
                x[k] = (x[k] & keep) | (0x54545454u & ~keep);
            }
        }
        uint32_t whi = 0u, wlo = 0u, wirr = 0u, pk = 0u;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t mA = __vcmpeq4(x[k], 0x41414141u), mC = __vcmpeq4(x[k], 0x43434343u), mG = __vcmpeq4(x[k], 0x47474747u),
                           mT = __vcmpeq4(x[k], 0x54545454u);
            const uint32_t code = 0x03030303u ^ ((mA & 0x03030303u) | (mC & 0x02020202u) | (mG & 0x01010101u)); // one 2-bit code per byte
            const uint32_t odd = ~(mA | mC | mG | mT) & 0x01010101u;
            whi |= ((((code >> 1) & 0x01010101u) * 0x01020408u) >> 24) << (4 * k); // byte q -> bit q (product bits 24..27, nothing above)
            wlo |= (((code & 0x01010101u) * 0x01020408u) >> 24) << (4 * k);
            wirr |= ((odd * 0x01020408u) >> 24) << (4 * k);
            pk |= ((code * 0x40100401u) >> 24) << (8 * k); // first base of a byte in bits 7:6; bytes little-endian in the word
        }
        const int64_t w = g >> 5;
        // the two granules of a 32-base word sit in neighbouring lanes: the even lane writes the plane words
        const uint32_t ohi = __shfl_down_sync(0xffffffffu, whi, 1), olo = __shfl_down_sync(0xffffffffu, wlo, 1),
                       oirr = __shfl_down_sync(0xffffffffu, wirr, 1);
        if (g < g_line) {
            packed[2 * w + (lane & 1)] = pk;
            if (!(lane & 1)) {
                hi[w] = whi | (ohi << 16);
                lo[w] = wlo | (olo << 16);
                irr[w] = wirr | (oirr << 16);
            }
            if (wirr) {
                bad += (unsigned long long)__popc(wirr);
                atomicOr(&flags[a], PB_FLAG_IRREGULAR);
            }
        }
    }
    if (bad) atomicAdd(nirr, bad);
}

// second pass, only when the set holds bytes outside {A,C,G,T}: list them as (line position, byte)
__global__ void __launch_bounds__(256)
collect_exceptions_kernel(IngestSrc src, const int64_t *__restrict__ base, const int32_t *__restrict__ len, int64_t n,
                          int64_t nwords, const uint32_t *__restrict__ irr, unsigned long long *__restrict__ cursor,
                          int64_t *__restrict__ exc_pos, uint8_t *__restrict__ exc_val, int32_t *__restrict__ exc_seq)
{
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= nwords) return;
    uint32_t m = irr[w];
    if (!m) return;
    int64_t a = 0, b = n;
    while (b - a > 1) {
        int64_t mid = (a + b) >> 1;
        if (base[mid] <= w * 32) a = mid; else b = mid;
    }
    const int64_t st = src.tstride ? (int64_t)src.tstride[a] : 1;
    while (m) {
        const int t = __ffs(m) - 1;
        m &= m - 1;
        const int64_t g = w * 32 + t, rel = g - base[a];
        const unsigned long long slot = atomicAdd(cursor, 1ull);
        exc_pos[slot] = g;
        exc_val[slot] = src.text[src.toff[a] + rel * st];
        exc_seq[slot] = (int32_t)a;
    }
}

int pb_seqset_build(pb_ctx *ctx, const void *d_text, const int64_t *h_toff, const int32_t *h_len,
                    const int32_t *h_stride, int64_t n, int src_mode, pb_seqset **out, int64_t text_bytes)
{
    pb_seqset *s = new pb_seqset();
    s->ctx = ctx;
    s->n = n;
    s->base.resize(n + 1);
    s->len.assign(h_len, h_len + n);
    int64_t g = 0;
    for (int64_t i = 0; i < n; ++i) {
        if (h_len[i] < 0) { delete s; return pb_fail(ctx, PB_ERR_ARG, "sequence %lld has negative length", (long long)i); }
        s->base[i] = g;
        g = (g + h_len[i] + 16 + 31) & ~(int64_t)31;
    }
    s->base[n] = g;
    s->total = ((g + 127) & ~(int64_t)127) + 128;
    const int64_t nw = s->nwords();
    int r;
    const bool htrace = getenv("PB_HOST_TRACE") != nullptr; // host wall time of this call's phases on stderr
    auto hnow = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double ht0 = htrace ? hnow() : 0.0;
    double ht1 = 0, ht2 = 0, ht3 = 0;
#define TRYS(x) do { r = (x); if (r != PB_OK) { delete s; return r; } } while (0)
    TRYS(s->d_base.alloc(ctx, (n + 1) * sizeof(int64_t)));
    TRYS(s->d_len.alloc(ctx, (n + 1) * sizeof(int32_t)));
    TRYS(s->d_flags.alloc_zero(ctx, (n + 1) * sizeof(uint32_t)));
    TRYS(s->d_hi.alloc(ctx, (nw + 4) * sizeof(uint32_t)));
    TRYS(s->d_lo.alloc(ctx, (nw + 4) * sizeof(uint32_t)));
    TRYS(s->d_packed.alloc(ctx, (2 * nw + 8) * sizeof(uint32_t)));
    TRYS(s->d_irr.alloc(ctx, (nw + 4) * sizeof(uint32_t)));
    DevBuf d_toff, d_stride, d_nirr;
    TRYS(d_nirr.alloc_zero(ctx, 16));
    TRYS(d_toff.alloc(ctx, (n + 1) * sizeof(int64_t)));
    const double hta = htrace ? hnow() : 0.0;
    TRYS(pb_h2d(ctx, s->d_base.p, s->base.data(), (n + 1) * sizeof(int64_t)));
    const double htb = htrace ? hnow() : 0.0;
    TRYS(pb_h2d(ctx, s->d_len.p, s->len.data(), n * sizeof(int32_t)));
    TRYS(pb_h2d(ctx, d_toff.p, h_toff, n * sizeof(int64_t)));

    if (h_stride) {
        TRYS(d_stride.alloc(ctx, (n + 1) * sizeof(int32_t)));
        TRYS(pb_h2d(ctx, d_stride.p, h_stride, n * sizeof(int32_t)));
    }
    if (htrace) ht1 = hnow();
    // guard words past the line read as code 3
    cudaMemsetAsync(s->d_hi.as<uint32_t>() + nw, 0xff, 4 * sizeof(uint32_t), ctx->stream);
    cudaMemsetAsync(s->d_lo.as<uint32_t>() + nw, 0xff, 4 * sizeof(uint32_t), ctx->stream);
    cudaMemsetAsync(s->d_packed.as<uint32_t>() + 2 * nw, 0xff, 8 * sizeof(uint32_t), ctx->stream);
    cudaMemsetAsync(s->d_irr.as<uint32_t>() + nw, 0, 4 * sizeof(uint32_t), ctx->stream);
    IngestSrc src;
    src.text = (const uint8_t *)d_text;
    src.toff = d_toff.as<int64_t>();
    src.tstride = h_stride ? d_stride.as<int32_t>() : nullptr;
    src.packed = src_mode;
    int64_t blocks = std::min<int64_t>(((nw + INGEST_RUN - 1) / INGEST_RUN + 7) / 8, (int64_t)ctx->sm_count * 16);
    if (blocks < 1) blocks = 1;
    bool forward = true; // every view reads its text left to right
    for (int64_t i = 0; h_stride && forward && i < n; ++i) forward = h_stride[i] == 1;
    pb_timer_begin(ctx, PB_T_INGEST);
    if (src_mode == PB_SRC_TEXT && forward && text_bytes >= 0 && n > 0 && (reinterpret_cast<uintptr_t>(d_text) & 3) == 0) {
        const int64_t warps = (nw * 32 + 512 * INGEST_STEPS - 1) / (512 * INGEST_STEPS);
        ingest_text_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, ctx->stream>>>(
            (const uint8_t *)d_text, text_bytes, d_toff.as<int64_t>(), s->d_base.as<int64_t>(), s->d_len.as<int32_t>(), n, nw,
            s->d_hi.as<uint32_t>(), s->d_lo.as<uint32_t>(), s->d_packed.as<uint32_t>(), s->d_irr.as<uint32_t>(),
            s->d_flags.as<uint32_t>(), d_nirr.as<unsigned long long>());
    } else
        ingest_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(src, s->d_base.as<int64_t>(), s->d_len.as<int32_t>(), n, nw,
                                                                 s->d_hi.as<uint32_t>(), s->d_lo.as<uint32_t>(),
                                                                 s->d_packed.as<uint32_t>(), s->d_irr.as<uint32_t>(),
                                                                 s->d_flags.as<uint32_t>(), d_nirr.as<unsigned long long>());
    pb_timer_end(ctx, PB_T_INGEST);
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { delete s; return pb_fail(ctx, PB_ERR_CUDA, "ingest launch failed: %s", cudaGetErrorString(e)); }
    s->flags.assign((size_t)n, 0u);
    unsigned long long nirr = 0;
    if (htrace) ht2 = hnow();
    TRYS(pb_d2h(ctx, &nirr, d_nirr.p, 8));
    TRYS(pb_sync(ctx)); // d_toff/d_stride are released after the kernel (stream-ordered) but the count is needed now
    if (htrace) ht3 = hnow();
    if (nirr > 0) { // the per-sequence flags only say something when a byte outside {A,C,G,T} was seen
        TRYS(pb_d2h(ctx, s->flags.data(), s->d_flags.p, n * sizeof(uint32_t)));
        TRYS(pb_sync(ctx));
    }
    s->tab.assign((size_t)n, 0x41414141u); // unused slots hold 'A': a value the exception list never contains
    s->tab_count.assign((size_t)n, 0);
    if (nirr > 0 && src_mode == PB_SRC_TEXT) {
        // rare path: list the offending bytes, sort them by line position on the host, build the per-sequence tables
        DevBuf d_cursor, d_pos, d_val, d_seq;
        TRYS(d_cursor.alloc_zero(ctx, 16));
        TRYS(d_pos.alloc(ctx, nirr * 8));
        TRYS(d_val.alloc(ctx, nirr + 16));
        TRYS(d_seq.alloc(ctx, nirr * 4));
        collect_exceptions_kernel<<<(unsigned)((nw + 255) / 256), 256, 0, ctx->stream>>>(
            src, s->d_base.as<int64_t>(), s->d_len.as<int32_t>(), n, nw, s->d_irr.as<uint32_t>(), d_cursor.as<unsigned long long>(),
            d_pos.as<int64_t>(), d_val.as<uint8_t>(), d_seq.as<int32_t>());
        ctx->launches++;
        std::vector<int64_t> pos((size_t)nirr);
        std::vector<uint8_t> val((size_t)nirr);
        std::vector<int32_t> seq((size_t)nirr);
        TRYS(pb_d2h(ctx, pos.data(), d_pos.p, nirr * 8));
        TRYS(pb_d2h(ctx, val.data(), d_val.p, nirr));
        TRYS(pb_d2h(ctx, seq.data(), d_seq.p, nirr * 4));
        TRYS(pb_sync(ctx));
        std::vector<size_t> order((size_t)nirr);
        for (size_t k = 0; k < order.size(); ++k) order[k] = k;
        std::sort(order.begin(), order.end(), [&](size_t x, size_t y) { return pos[x] < pos[y]; });
        std::vector<int64_t> spos((size_t)nirr);
        std::vector<uint8_t> sval((size_t)nirr);
        for (size_t k = 0; k < order.size(); ++k) {
            spos[k] = pos[order[k]];
            sval[k] = val[order[k]];
            const int32_t q = seq[order[k]];
            uint8_t &cnt = s->tab_count[q];
            if (cnt == 255) continue;
            bool seen = false;
            for (int t = 0; t < cnt; ++t) seen = seen || ((s->tab[q] >> (8 * t)) & 0xFF) == sval[k];
            if (!seen) {
                if (cnt == 4) cnt = 255;
                else { s->tab[q] = (s->tab[q] & ~(0xFFu << (8 * cnt))) | ((uint32_t)sval[k] << (8 * cnt)); ++cnt; }
            }
        }
        s->nexc = (int64_t)nirr;
        TRYS(s->d_exc_pos.alloc(ctx, nirr * 8));
        TRYS(s->d_exc_val.alloc(ctx, nirr + 16));
        TRYS(pb_h2d(ctx, s->d_exc_pos.p, spos.data(), nirr * 8));
        TRYS(pb_h2d(ctx, s->d_exc_val.p, sval.data(), nirr));
        TRYS(pb_sync(ctx));
    }
    TRYS(s->d_tab.alloc(ctx, (size_t)(n + 1) * 4));
    TRYS(pb_h2d(ctx, s->d_tab.p, s->tab.data(), (size_t)n * 4));
    TRYS(pb_sync(ctx));
    if (htrace)
        fprintf(stderr, "[pb_host_trace] pb_seqset_build (ms): allocs %.2f first table copy %.2f other table copies %.2f launch %.2f first sync %.2f tables+sync %.2f\n",
                hta - ht0, htb - hta, ht1 - htb, ht2 - ht1, ht3 - ht2, hnow() - ht3);
#undef TRYS
    *out = s;
    return PB_OK;
}

// text blob extent touched by the views: [lo, hi)
static bool text_extent(const int64_t *off, const int32_t *len, const int32_t *stride, int64_t n, int64_t *lo, int64_t *hi)
{
    int64_t a = INT64_MAX, b = INT64_MIN;
    for (int64_t i = 0; i < n; ++i) {
        if (len[i] <= 0) continue;
        int64_t st = stride ? stride[i] : 1;
        if (st != 1 && st != -1) return false;
        int64_t first = off[i], last = off[i] + (int64_t)(len[i] - 1) * st;
        a = std::min(a, std::min(first, last));
        b = std::max(b, std::max(first, last) + 1);
    }
    if (a > b) { a = 0; b = 0; }
    *lo = a; *hi = b;
    return true;
}

extern "C" int pb_seqset_from_text(pb_ctx *ctx, const char *text, const int64_t *off, const int32_t *len,
                                   const int32_t *stride, int64_t n, pb_seqset **out)
{
    if (!ctx || !out || n < 0 || (n > 0 && (!text || !off || !len))) return pb_fail(ctx, PB_ERR_ARG, "pb_seqset_from_text: bad argument");
    *out = nullptr;
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    int64_t lo, hi;
    if (!text_extent(off, len, stride, n, &lo, &hi)) return pb_fail(ctx, PB_ERR_ARG, "stride must be +1 or -1");
    DevBuf d_text;
    PB_TRY(d_text.alloc(ctx, (size_t)(hi - lo) + 16));
    pb_timer_begin(ctx, PB_T_H2D);
    PB_TRY(pb_h2d(ctx, d_text.p, text + lo, (size_t)(hi - lo)));
    pb_timer_end(ctx, PB_T_H2D);
    std::vector<int64_t> rel(off, off + n);
    for (auto &x : rel) x -= lo;
    return pb_seqset_build(ctx, d_text.p, rel.data(), len, stride, n, PB_SRC_TEXT, out, (hi - lo) + 16);
}

extern "C" int pb_seqset_from_device_text(pb_ctx *ctx, const void *d_text, size_t text_bytes, const int64_t *off,
                                          const int32_t *len, const int32_t *stride, int64_t n, pb_seqset **out)
{
    if (!ctx || !out || n < 0 || (n > 0 && (!d_text || !off || !len))) return pb_fail(ctx, PB_ERR_ARG, "pb_seqset_from_device_text: bad argument");
    *out = nullptr;
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    int64_t lo, hi;
    if (!text_extent(off, len, stride, n, &lo, &hi)) return pb_fail(ctx, PB_ERR_ARG, "stride must be +1 or -1");
    if (lo < 0 || (size_t)hi > text_bytes) return pb_fail(ctx, PB_ERR_ARG, "views reach outside the device text blob");
    return pb_seqset_build(ctx, d_text, off, len, stride, n, PB_SRC_TEXT, out, (int64_t)text_bytes);
}

extern "C" int pb_seqset_from_bin(pb_ctx *ctx, const uint8_t *bin, size_t nbytes, int min_excl, int max_excl,
                                  pb_seqset **out)
{
    if (!ctx || !out || (!bin && nbytes)) return pb_fail(ctx, PB_ERR_ARG, "pb_seqset_from_bin: bad argument");
    *out = nullptr;
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    // record walk, spaced_seed.cpp:330-342: u32 length, ceil(len/4) body bytes, records back to back
    std::vector<int64_t> off;
    std::vector<int32_t> len;
    size_t p = 0;
    while (p + 4 <= nbytes) {
        uint32_t l;
        memcpy(&l, bin + p, 4);
        size_t body = ((size_t)l + 3) / 4;
        if (p + 4 + body > nbytes) return pb_fail(ctx, PB_ERR_ARG, "truncated .bin record at byte %zu", p);
        if ((int64_t)l > min_excl && (int64_t)l < max_excl) {
            off.push_back((int64_t)p + 4);
            len.push_back((int32_t)l);
        }
        p += 4 + body;
    }
    DevBuf d_bin;
    PB_TRY(d_bin.alloc(ctx, nbytes + 16));
    pb_timer_begin(ctx, PB_T_H2D);
    PB_TRY(pb_h2d(ctx, d_bin.p, bin, nbytes));
    pb_timer_end(ctx, PB_T_H2D);
    PB_TRY(pb_seqset_build(ctx, d_bin.p, off.data(), len.data(), nullptr, (int64_t)off.size(), PB_SRC_PACKED, out));
    // keep the image: dna_seq::seed_at's pos%4==0 branch reads raw bytes of it, possibly of later records (SURVEY Q-S1)
    pb_seqset *s = *out;
    s->image_bytes = (int64_t)nbytes;
    s->d_image.p = d_bin.p; s->d_image.bytes = d_bin.bytes; s->d_image.ctx = ctx;
    d_bin.p = nullptr; d_bin.bytes = 0;
    int r = s->d_recoff.alloc(ctx, (off.size() + 1) * sizeof(int64_t));
    if (r == PB_OK) r = pb_h2d(ctx, s->d_recoff.p, off.data(), off.size() * sizeof(int64_t));
    if (r == PB_OK) r = pb_sync(ctx);
    if (r != PB_OK) { pb_seqset_free(s); *out = nullptr; }
    return r;
}

int pb_seqset_reversed(pb_ctx *ctx, const pb_seqset *s, pb_seqset **out)
{ // every sequence read backwards (seq_accessor with forward == false, dna_seq.h:185-233), same line layout
    for (int64_t i = 0; i < s->n; ++i)
        if (s->flags[i] & PB_FLAG_IRREGULAR)
            return pb_fail(ctx, PB_ERR_ALPHABET, "backward views of sequences with bytes outside {A,C,G,T} are not supported");
    return pb_seqset_build(ctx, s->d_packed.p, s->base.data(), s->len.data(), nullptr, s->n, PB_SRC_REVLINE, out);
}

extern "C" void pb_seqset_free(pb_seqset *s)
{
    if (!s) return;
    cudaSetDevice(s->ctx->device);
    delete s;
}

extern "C" int64_t pb_seqset_count(const pb_seqset *s) { return s ? s->n : 0; }
extern "C" int32_t pb_seqset_length(const pb_seqset *s, int64_t i) { return (s && i >= 0 && i < s->n) ? s->len[i] : -1; }

// ---- decode back to text (bin2text, dna_seq.h:133-145) and raw packed access --------------------------------

__global__ void unpack_text_kernel(const uint8_t *__restrict__ packed, int64_t first_base, int64_t count, char *out)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    int64_t g = first_base + i;
    uint32_t code = (packed[g >> 2] >> (6 - 2 * (g & 3))) & 3u;
    out[i] = code == 0 ? 'A' : code == 1 ? 'C' : code == 2 ? 'G' : 'T';
}

extern "C" int pb_seqset_text(pb_ctx *ctx, const pb_seqset *s, int64_t i, char *out, size_t cap)
{
    if (!ctx || !s || !out || i < 0 || i >= s->n) return pb_fail(ctx, PB_ERR_ARG, "pb_seqset_text: bad argument");
    const int64_t L = s->len[i];
    if (cap <= (size_t)L) return pb_fail(ctx, PB_ERR_ARG, "pb_seqset_text: cap %zu <= length %lld", cap, (long long)L);
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    DevBuf d;
    PB_TRY(d.alloc(ctx, (size_t)L + 16));
    if (L > 0) {
        unpack_text_kernel<<<(unsigned)((L + 255) / 256), 256, 0, ctx->stream>>>(s->d_packed.as<uint8_t>(), s->base[i], L, d.as<char>());
        PB_LAUNCH_CHECK(ctx);
        PB_TRY(pb_d2h(ctx, out, d.p, (size_t)L));
    }
    PB_TRY(pb_sync(ctx));
    out[L] = '\0';
    return PB_OK;
}

__global__ void mask_tail_kernel(uint8_t *body, int64_t len)
{ // text2bin zero-pads the last byte (dna_seq.h:147-159); the padded line holds code 3 there
    if (threadIdx.x == 0 && blockIdx.x == 0 && (len & 3)) body[len >> 2] &= (uint8_t)(0xFF << (8 - 2 * (len & 3)));
}

extern "C" int pb_seqset_packed(pb_ctx *ctx, const pb_seqset *s, int64_t i, uint8_t *out, size_t cap)
{
    if (!ctx || !s || !out || i < 0 || i >= s->n) return pb_fail(ctx, PB_ERR_ARG, "pb_seqset_packed: bad argument");
    const int64_t L = s->len[i];
    const size_t nb = (size_t)(L + 3) / 4;
    if (cap < nb) return pb_fail(ctx, PB_ERR_ARG, "pb_seqset_packed: cap too small");
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    DevBuf d;
    PB_TRY(d.alloc(ctx, nb + 16));
    if (nb) {
        PB_CUDA(ctx, cudaMemcpyAsync(d.p, s->d_packed.as<uint8_t>() + (s->base[i] >> 2), nb, cudaMemcpyDeviceToDevice, ctx->stream));
        mask_tail_kernel<<<1, 32, 0, ctx->stream>>>(d.as<uint8_t>(), L);
        PB_LAUNCH_CHECK(ctx);
        PB_TRY(pb_d2h(ctx, out, d.p, nb));
    }
    return pb_sync(ctx);
}

// ---- dna_seq statics as batched device calls -------------------------------------------------------------

extern "C" int pb_text2bin(pb_ctx *ctx, const char *text, size_t tlen, uint8_t *out, size_t cap, size_t *written)
{
    if (!ctx || (!text && tlen) || !out) return pb_fail(ctx, PB_ERR_ARG, "pb_text2bin: bad argument");
    const size_t blen = 4 + (tlen + 3) / 4;
    if (cap < blen) return pb_fail(ctx, PB_ERR_ARG, "pb_text2bin: buffer of %zu bytes < record of %zu (dna_seq.h:118 asserts)", cap, blen);
    int64_t off = 0;
    int32_t len = (int32_t)tlen;
    pb_seqset *s = nullptr;
    PB_TRY(pb_seqset_from_text(ctx, text ? text : "", &off, &len, nullptr, 1, &s));
    uint32_t l32 = (uint32_t)tlen;
    memcpy(out, &l32, 4);
    int r = pb_seqset_packed(ctx, s, 0, out + 4, cap - 4);
    pb_seqset_free(s);
    if (written) *written = blen;
    return r;
}

extern "C" int pb_bin2text(pb_ctx *ctx, const uint8_t *rec, char *out, size_t cap, size_t *tlen)
{
    if (!ctx || !rec || !out) return pb_fail(ctx, PB_ERR_ARG, "pb_bin2text: bad argument");
    uint32_t l;
    memcpy(&l, rec, 4);
    if (cap <= l) return pb_fail(ctx, PB_ERR_ARG, "pb_bin2text: buffer of %zu bytes <= length %u (dna_seq.h:138 asserts)", cap, l);
    pb_seqset *s = nullptr;
    PB_TRY(pb_seqset_from_bin(ctx, rec, 4 + ((size_t)l + 3) / 4, -1, INT32_MAX, &s));
    int r = pb_seqset_text(ctx, s, 0, out, cap);
    pb_seqset_free(s);
    if (tlen) *tlen = l;
    return r;
}

__global__ void decode_kernel(const uint32_t *__restrict__ codes, int64_t n, char *__restrict__ out)
{ // dna_seq::decode, dna_seq.h:101-107: byte k of the word holds bases 4k..4k+3, first base in bits 7:6
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * 16) return;
    uint32_t w = codes[i >> 4];
    int k = (int)(i & 15);
    uint32_t code = (w >> (8 * (k >> 2) + 6 - 2 * (k & 3))) & 3u;
    out[i] = code == 0 ? 'A' : code == 1 ? 'C' : code == 2 ? 'G' : 'T';
}

extern "C" int pb_decode_batch(pb_ctx *ctx, const uint32_t *codes, int64_t n, char *out16)
{
    if (!ctx || n < 0 || (n && (!codes || !out16))) return pb_fail(ctx, PB_ERR_ARG, "pb_decode_batch: bad argument");
    if (!n) return PB_OK;
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    DevBuf dc, dout;
    PB_TRY(dc.alloc(ctx, n * 4));
    PB_TRY(dout.alloc(ctx, n * 16));
    PB_TRY(pb_h2d(ctx, dc.p, codes, n * 4));
    decode_kernel<<<(unsigned)((n * 16 + 255) / 256), 256, 0, ctx->stream>>>(dc.as<uint32_t>(), n, dout.as<char>());
    PB_LAUNCH_CHECK(ctx);
    PB_TRY(pb_d2h(ctx, out16, dout.p, n * 16));
    return pb_sync(ctx);
}

__global__ void encode_kernel(const uint8_t *__restrict__ text, int64_t text_len, const int64_t *__restrict__ off, int64_t n,
                              uint32_t *__restrict__ out)
{ // dna_seq::encode, dna_seq.h:86-96
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int64_t o = off[i];
    uint32_t w = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        int64_t p = o + k;
        uint32_t ch = (p >= 0 && p < text_len) ? text[p] : 0u;
        w |= c2i(ch) << (8 * (k >> 2) + 6 - 2 * (k & 3));
    }
    out[i] = w;
}

extern "C" int pb_encode_batch(pb_ctx *ctx, const char *text, size_t text_len, const int64_t *off, int64_t n, uint32_t *out)
{
    if (!ctx || n < 0 || (n && (!off || !out)) || (!text && text_len)) return pb_fail(ctx, PB_ERR_ARG, "pb_encode_batch: bad argument");
    if (!n) return PB_OK;
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    DevBuf dt, doff, dout;
    PB_TRY(dt.alloc(ctx, text_len + 16));
    PB_TRY(doff.alloc(ctx, n * 8));
    PB_TRY(dout.alloc(ctx, n * 4));
    PB_TRY(pb_h2d(ctx, dt.p, text, text_len));
    PB_TRY(pb_h2d(ctx, doff.p, off, n * 8));
    encode_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(dt.as<uint8_t>(), (int64_t)text_len, doff.as<int64_t>(), n, dout.as<uint32_t>());
    PB_LAUNCH_CHECK(ctx);
    PB_TRY(pb_d2h(ctx, out, dout.p, n * 4));
    return pb_sync(ctx);
}

__global__ void seed_at_kernel(const uint8_t *__restrict__ rec, int64_t rec_bytes, const int32_t *__restrict__ pos, int64_t n,
                               int quirk, uint32_t *__restrict__ out)
{ // dna_seq::seed_at, dna_seq.h:62-76
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int p = pos[i];
    auto byte_at = [&](int64_t o) -> uint32_t { return (o >= 0 && o < rec_bytes) ? rec[o] : 0u; };
    uint32_t w = 0;
    if (quirk && (p & 3) == 0) { // Q-S1: *((unsigned*)(pbin + 4 + pos))
        for (int k = 0; k < 4; ++k) w |= byte_at(4 + (int64_t)p + k) << (8 * k);
    } else {
        const int64_t b0 = 4 + (p >> 2);
        const unsigned ls = (p & 3) << 1, rs = 8 - ls;
        for (int k = 0; k < 4; ++k) {
            uint32_t v = ((byte_at(b0 + k) << ls) | (ls ? (byte_at(b0 + k + 1) >> rs) : 0u)) & 0xFFu;
            w |= v << (8 * k);
        }
    }
    out[i] = w;
}

extern "C" int pb_seed_at_batch(pb_ctx *ctx, const uint8_t *rec, size_t rec_bytes, const int32_t *pos, int64_t n, int quirk,
                                uint32_t *out)
{
    if (!ctx || n < 0 || (n && (!rec || !pos || !out))) return pb_fail(ctx, PB_ERR_ARG, "pb_seed_at_batch: bad argument");
    if (!n) return PB_OK;
    PB_CUDA(ctx, cudaSetDevice(ctx->device));
    DevBuf dr, dp, dout;
    PB_TRY(dr.alloc(ctx, rec_bytes + 16));
    PB_TRY(dp.alloc(ctx, n * 4));
    PB_TRY(dout.alloc(ctx, n * 4));
    PB_TRY(pb_h2d(ctx, dr.p, rec, rec_bytes));
    PB_TRY(pb_h2d(ctx, dp.p, pos, n * 4));
    seed_at_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(dr.as<uint8_t>(), (int64_t)rec_bytes, dp.as<int32_t>(), n, quirk, dout.as<uint32_t>());
    PB_LAUNCH_CHECK(ctx);
    PB_TRY(pb_d2h(ctx, out, dout.p, n * 4));
    return pb_sync(ctx);
}

extern "C" uint32_t pb_parse_pattern(const char *pattern)
{ // spaced_seed.cpp:166-180: '1' -> 'T' (bits 11), anything else -> 'A' (00), padded with A to 16, truncated at 16
    uint32_t w = 0;
    if (!pattern) return 0;
    size_t len = strlen(pattern);
    if (len > 16) len = 16;
    for (size_t k = 0; k < len; ++k)
        if (pattern[k] == '1') w |= 3u << (8 * (k >> 2) + 6 - 2 * (k & 3));
    return w;
}
