// ac3_encode_host.inl - host side of the encoder: tables, context, the C ABI of
// include/ac3enc_batch.h and the drop-in AC3_encode_* API of include/ac3enc.h.
// Included at the end of ac3_encode.cu.

namespace ac3e {

static const int h_freqs[3] = {48000, 44100, 32000};

static int16_t h_fix15(float a)             // ac3enc.cpp:427-439
{
    int v = (int)(a * (float)(1 << 15));
    if (v < -32767) v = -32767; else if (v > 32767) v = 32767;
    return (int16_t)v;
}

static void build_enc_tables(EncTables* T)
{
    memset(T, 0, sizeof(*T));
    // KBD (alpha = 5) window in Q15, truncated (ac3tab.h:14-47 holds the same values as literals)
    {
        double sum = 0, cum[256];
        for (int i = 0; i < 256; i++) {
            double x = i * (256 - i) * (5 * M_PI / 256) * (5 * M_PI / 256), b = 1;
            for (int k = 100; k > 0; k--) b = b * x / (k * k) + 1;
            sum += b;
            cum[i] = sum;
        }
        sum++;
        for (int i = 0; i < 256; i++) {
            int v = (int)floor(sqrt(cum[i] / sum) * 32768.0);
            T->window[i] = (int16_t)(v > 32767 ? 32767 : v);
        }
    }
    // fft / mdct twiddles, float arithmetic exactly as the reference writes it (ac3enc.cpp:441-459, 1098-1102)
    for (int i = 0; i < 64; i++) {
        float alpha = (float)(2 * M_PI * (float)i / (float)128);
        T->costab[i] = h_fix15((float)cos(alpha));
        T->sintab[i] = h_fix15((float)sin(alpha));
    }
    for (int i = 0; i < 128; i++) {
        int m = 0;
        for (int j = 0; j < 7; j++) m |= ((i >> j) & 1) << (6 - j);
        T->rev[i] = (uint8_t)m;
        float alpha = (float)(2 * M_PI * (i + 1.0 / 8.0) / (float)512);
        T->xcos1[i] = h_fix15((float)-cos(alpha));
        T->xsin1[i] = h_fix15((float)-sin(alpha));
    }
    for (int i = 0; i < 256; i++) {          // ac3enc.cpp:998-1016
        unsigned c = (unsigned)i << 8;
        for (int j = 0; j < 8; j++) c = (c & 0x8000) ? (((c << 1) & 0xffff) ^ 0x8005) : (c << 1);
        T->crc_table[i] = (uint16_t)c;
        T->masktab[i] = ac3_masktab[i];
        T->latab[i] = ac3_latab[i];
    }
    for (int i = 0; i < 150; i++) T->hth[i] = ac3_hth[i];
    for (int i = 0; i < 64; i++) T->baptab[i] = ac3_baptab[i];
    for (int i = 0; i < 51; i++) T->bndtab[i] = ac3_bndtab[i];
    T->bndtab[51] = 253;
    for (int b = 0; b < 16; b++) {
        T->width[b] = ac3_bap_bits[b];
        T->plain_bits[b] = (b == 0 || b == 1 || b == 2 || b == 4) ? 0 : ac3_bap_bits[b];
        // counters of the packer and the search: bits of an ungrouped field, one count per grouped class
        T->taba[b] = T->plain_bits[b] | (uint32_t)(b == 1) << 8 | (uint32_t)(b == 2) << 16 | (uint32_t)(b == 4) << 24;
        // quantiser per bap (ac3enc.cpp:1363-1460): 3 / 5 / 7 / 11 / 15 levels symmetric, the rest asymmetric
        static const uint8_t levels[16] = {0, 3, 5, 7, 11, 15, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        const uint32_t qbits = b < 6 ? 1 : b == 14 ? 14 : b == 15 ? 16 : b - 1;
        const uint32_t cls = b == 1 ? 1 : b == 2 ? 2 : b == 4 ? 3 : 0;
        T->tabq[b].x = b ? (levels[b] | qbits << 8 | (uint32_t)ac3_bap_bits[b] << 16 | cls << 24) : 0u;
        T->tabq[b].y = cls ? 1u << (10 * (cls - 1)) : 0u;
    }
    for (int a = 0; a < 64; a++) T->tabc[a] = T->taba[ac3_baptab[a]];
}

struct EncConfig {
    int nch_all, nch, lfe, acmod, fscod, halfrate, bsid, frmsizecod, frame_words;
};

static bool enc_config(int freq, int bitrate, int channels, EncConfig* c)    // ac3enc.cpp:1019-1077
{
    static const uint8_t acmod_defs[6] = {1, 2, 3, 6, 7, 7};
    if (channels < 1 || channels > 6) return false;
    c->acmod = acmod_defs[channels - 1];
    c->lfe = channels == 6;
    c->nch_all = channels;
    c->nch = channels > 5 ? 5 : channels;
    bool found = false;
    for (int i = 0; i < 3 && !found; i++)
        for (int j = 0; j < 3; j++)
            if ((h_freqs[j] >> i) == freq) { c->halfrate = i; c->fscod = j; found = true; break; }
    if (!found) return false;
    c->bsid = 8 + c->halfrate;
    bitrate /= 1000;
    int i;
    for (i = 0; i < 19; i++)
        if ((ac3_bitrate_kbps[i] >> c->halfrate) == bitrate) break;
    if (i == 19) return false;
    c->frmsizecod = i << 1;
    c->frame_words = (bitrate * 1000 * 1536) / (freq * 16);
    return c->frame_words > 0;
}

static unsigned h_mul_poly(unsigned a, unsigned b, unsigned poly)
{
    unsigned c = 0;
    while (a) { if (a & 1) c ^= b; a >>= 1; b <<= 1; if (b & 0x10000) b ^= poly; }
    return c;
}
static unsigned h_pow_poly(unsigned a, unsigned n, unsigned poly)
{
    unsigned r = 1;
    while (n) { if (n & 1) r = h_mul_poly(r, a, poly); a = h_mul_poly(a, a, poly); n >>= 1; }
    return r;
}

}  // namespace ac3e

struct ac3_batch_s {
    int device = 0;
    int num_sms = 0;
    int* d_counter = nullptr;
    char err[256] = {0};
    long launches = 0;
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used = 0;
    struct Buf { void* p = nullptr; size_t cap = 0; } b_pcm, b_out, b_status, b_carry, b_coef, b_shift, b_strat, b_enc, b_bap, b_snr, b_done, b_scarry;
    int slice_frames = 32;         // frames per work unit (AC3_B200_SLICE_FRAMES)
    // host-pointer calls: PCM H2D | kernel | frames D2H on three streams, chunks of streams
    cudaStream_t s_in = nullptr, s_run = nullptr, s_out = nullptr;
    cudaEvent_t ev_in[16] = {}, ev_run[16] = {};
};

#define AC3_CUDA(call)                                                                       \
    do {                                                                                     \
        cudaError_t e_ = (call);                                                             \
        if (e_ != cudaSuccess) {                                                             \
            snprintf(ctx->err, sizeof(ctx->err), "%s failed: %s", #call, cudaGetErrorString(e_)); \
            return -1;                                                                       \
        }                                                                                    \
    } while (0)

static int ac3_ensure(ac3_batch_t* ctx, ac3_batch_s::Buf& b, size_t bytes)
{
    if (bytes <= b.cap) return 0;
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
    size_t cap = bytes + bytes / 8 + 256;
    AC3_CUDA(cudaMalloc(&b.p, cap));
    b.cap = cap;
    return 0;
}

#pragma GCC visibility push(default)
extern "C" {

ac3_batch_t* ac3_batch_create(int device)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return nullptr;
    if (cudaSetDevice(device) != cudaSuccess) return nullptr;
    ac3_batch_t* ctx = new (std::nothrow) ac3_batch_s();
    if (!ctx) return nullptr;
    ctx->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete ctx; return nullptr; }
    ctx->num_sms = prop.multiProcessorCount;
    const char* sf = getenv("AC3_B200_SLICE_FRAMES");
    if (sf && atoi(sf) > 0) ctx->slice_frames = atoi(sf);
    ac3e::EncTables* T = new ac3e::EncTables;
    ac3e::build_enc_tables(T);
    bool ok = cudaMemcpyToSymbol(ac3e::g_enc_tables, T, sizeof(*T)) == cudaSuccess;
    delete T;
    ok = ok && cudaMalloc(&ctx->d_counter, 16 * sizeof(int)) == cudaSuccess;
    ok = ok && cudaFuncSetAttribute(ac3e::ac3_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)(sizeof(ac3e::EncShared) + sizeof(ac3e::EncTables) + 32)) == cudaSuccess;
    if (!ok) { ac3_batch_destroy(ctx); return nullptr; }
    return ctx;
}

void ac3_batch_destroy(ac3_batch_t* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    ac3_batch_s::Buf* bufs[] = {&ctx->b_pcm, &ctx->b_out, &ctx->b_status, &ctx->b_carry, &ctx->b_coef,
                                &ctx->b_shift, &ctx->b_strat, &ctx->b_enc, &ctx->b_bap, &ctx->b_snr, &ctx->b_done,
                                &ctx->b_scarry};
    for (auto* b : bufs)
        if (b->p) cudaFree(b->p);
    if (ctx->d_counter) cudaFree(ctx->d_counter);
    for (auto e : ctx->ev_pool) cudaEventDestroy(e);
    if (ctx->s_in) {
        cudaStreamDestroy(ctx->s_in);
        cudaStreamDestroy(ctx->s_run);
        cudaStreamDestroy(ctx->s_out);
        for (int i = 0; i < 16; i++) { cudaEventDestroy(ctx->ev_in[i]); cudaEventDestroy(ctx->ev_run[i]); }
    }
    delete ctx;
}

const char* ac3_batch_last_error(ac3_batch_t* ctx) { return ctx ? ctx->err : "no context"; }

int ac3_batch_frame_bytes(int freq, int bitrate, int channels)
{
    ac3e::EncConfig c;
    return ac3e::enc_config(freq, bitrate, channels, &c) ? c.frame_words * 2 : 0;
}

long ac3_batch_launch_count(ac3_batch_t* ctx) { return ctx->launches; }

double ac3_batch_kernel_ms(ac3_batch_t* ctx, int* nlaunches)
{
    double total = 0;
    int n = 0;
    cudaSetDevice(ctx->device);
    for (size_t i = 0; i + 1 < ctx->ev_used; i += 2) {
        float ms = 0;
        cudaEventSynchronize(ctx->ev_pool[i + 1]);
        if (cudaEventElapsedTime(&ms, ctx->ev_pool[i], ctx->ev_pool[i + 1]) == cudaSuccess) { total += ms; n++; }
    }
    ctx->ev_used = 0;
    if (nlaunches) *nlaunches = n;
    return n ? total / n : 0.0;
}

int ac3_batch_encode(ac3_batch_t* ctx, const int16_t* pcm, int nstreams, int nframes, int freq, int bitrate,
                     int channels, const uint8_t* chmap, uint8_t* out, int32_t* status, ac3_stream_carry_t* carry,
                     const ac3_batch_debug_t* debug, int mem_flags, void* cuda_stream)
{
    using namespace ac3e;
    if (!ctx) return -1;
    ctx->err[0] = 0;
    EncConfig c;
    if (nstreams < 0 || nframes < 0 || !enc_config(freq, bitrate, channels, &c)) {
        snprintf(ctx->err, sizeof(ctx->err), "bad argument (configuration rejected as by AC3_encode_init)");
        return -3;
    }
    if (nstreams == 0 || nframes == 0) return 0;
    AC3_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const size_t total = (size_t)nstreams * nframes;
    const size_t pcm_bytes = total * 1536 * channels * 2, out_bytes = total * c.frame_words * 2;

    EncParams P;
    memset(&P, 0, sizeof(P));
    P.nstreams = nstreams;
    P.nframes = nframes;
    P.nch_all = c.nch_all; P.nch = c.nch; P.lfe = c.lfe; P.acmod = c.acmod; P.fscod = c.fscod;
    P.halfrate = c.halfrate; P.bsid = c.bsid; P.frmsizecod = c.frmsizecod; P.frame_words = c.frame_words;
    const int fs58 = (c.frame_words >> 1) + (c.frame_words >> 3);
    P.crc_inv = h_pow_poly(0x18005 >> 1, (unsigned)(16 * fs58 - 16), 0x18005);
    for (int i = 0; i < 6; i++) P.chmap[i] = chmap ? chmap[i < channels ? i : 0] : (uint8_t)i;
    for (int i = 0; i < channels; i++)
        if (P.chmap[i] >= channels) {
            snprintf(ctx->err, sizeof(ctx->err), "bad channel map");
            return -3;
        }
    P.work_counter = ctx->d_counter;
    const bool dev = (mem_flags & AC3_BATCH_DEVICE_PTRS) != 0;
    bool pipelined = false;
    if (dev) {
        P.pcm = pcm; P.out = out; P.status = status; P.carry = (EncCarry*)carry;
        if (debug) {
            P.dbg_coef = debug->coef; P.dbg_shift = debug->exp_shift; P.dbg_strategy = debug->strategy;
            P.dbg_enc = debug->encoded_exp; P.dbg_bap = debug->bap; P.dbg_snr = debug->snr;
        }
    } else {
        if (ac3_ensure(ctx, ctx->b_pcm, pcm_bytes) || ac3_ensure(ctx, ctx->b_out, out_bytes) ||
            ac3_ensure(ctx, ctx->b_status, total * 4)) return -1;
        pipelined = !st && !debug && nstreams >= 1024;
        if (!pipelined) AC3_CUDA(cudaMemcpyAsync(ctx->b_pcm.p, pcm, pcm_bytes, cudaMemcpyHostToDevice, st));
        P.pcm = (const int16_t*)ctx->b_pcm.p;
        P.out = (uint8_t*)ctx->b_out.p;
        P.status = (int32_t*)ctx->b_status.p;
        if (carry) {
            if (ac3_ensure(ctx, ctx->b_carry, sizeof(EncCarry) * (size_t)nstreams)) return -1;
            if (!pipelined)
                AC3_CUDA(cudaMemcpyAsync(ctx->b_carry.p, carry, sizeof(EncCarry) * (size_t)nstreams, cudaMemcpyHostToDevice, st));
            P.carry = (EncCarry*)ctx->b_carry.p;
        }
        if (debug && debug->coef) {
            if (ac3_ensure(ctx, ctx->b_coef, total * 9216 * 4) || ac3_ensure(ctx, ctx->b_shift, total * 36) ||
                ac3_ensure(ctx, ctx->b_strat, total * 36) || ac3_ensure(ctx, ctx->b_enc, total * 9216) ||
                ac3_ensure(ctx, ctx->b_bap, total * 9216) || ac3_ensure(ctx, ctx->b_snr, total * 8)) return -1;
            P.dbg_coef = (int32_t*)ctx->b_coef.p; P.dbg_shift = (int8_t*)ctx->b_shift.p;
            P.dbg_strategy = (uint8_t*)ctx->b_strat.p; P.dbg_enc = (uint8_t*)ctx->b_enc.p;
            P.dbg_bap = (uint8_t*)ctx->b_bap.p; P.dbg_snr = (int32_t*)ctx->b_snr.p;
        }
    }
    // all six dump pointers travel together
    if (!(P.dbg_coef && P.dbg_shift && P.dbg_strategy && P.dbg_enc && P.dbg_bap && P.dbg_snr))
        P.dbg_coef = nullptr, P.dbg_shift = nullptr, P.dbg_strategy = nullptr, P.dbg_enc = nullptr,
        P.dbg_bap = nullptr, P.dbg_snr = nullptr;

    const size_t smem = ((sizeof(EncTables) + 15) & ~(size_t)15) + sizeof(EncShared);
    int occ = 0;
    AC3_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, ac3_encode_kernel, kThreads, smem));
    if (occ < 1) {
        snprintf(ctx->err, sizeof(ctx->err), "encode kernel does not fit: %zu bytes of shared memory", smem);
        return -2;
    }
    int grid = ctx->num_sms * occ;
    if (grid > nstreams) grid = nstreams;
    // work units: slices of streams (see the kernel); the carry record links the slices of a stream
    P.slice_frames = ctx->slice_frames;
    P.nslices = 1;
    if (nframes > ctx->slice_frames && !P.dbg_coef) {
        P.nslices = (nframes + ctx->slice_frames - 1) / ctx->slice_frames;
        cudaStream_t s0 = pipelined ? ctx->s_run : st;
        if (pipelined && !ctx->s_in) {
            AC3_CUDA(cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking));
            AC3_CUDA(cudaStreamCreateWithFlags(&ctx->s_run, cudaStreamNonBlocking));
            AC3_CUDA(cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking));
            for (int i = 0; i < 16; i++) {
                AC3_CUDA(cudaEventCreateWithFlags(&ctx->ev_in[i], cudaEventDisableTiming));
                AC3_CUDA(cudaEventCreateWithFlags(&ctx->ev_run[i], cudaEventDisableTiming));
            }
            s0 = ctx->s_run;
        }
        if (ac3_ensure(ctx, ctx->b_done, (size_t)nstreams * sizeof(int))) return -1;
        AC3_CUDA(cudaMemsetAsync(ctx->b_done.p, 0, (size_t)nstreams * sizeof(int), s0));
        P.slice_done = (int*)ctx->b_done.p;
        if (!P.carry) {
            if (ac3_ensure(ctx, ctx->b_scarry, (size_t)nstreams * sizeof(EncCarry))) return -1;
            AC3_CUDA(cudaMemsetAsync(ctx->b_scarry.p, 0, (size_t)nstreams * sizeof(EncCarry), s0));
            P.carry = (EncCarry*)ctx->b_scarry.p;
        }
    }
    if (pipelined) {
        // Large host-pointer batches: chunks of streams flow through PCM H2D | kernel | frames D2H on three
        // CUDA streams (a chunk still fills every SM several times over).
        if (!ctx->s_in) {
            AC3_CUDA(cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking));
            AC3_CUDA(cudaStreamCreateWithFlags(&ctx->s_run, cudaStreamNonBlocking));
            AC3_CUDA(cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking));
            for (int i = 0; i < 16; i++) {
                AC3_CUDA(cudaEventCreateWithFlags(&ctx->ev_in[i], cudaEventDisableTiming));
                AC3_CUDA(cudaEventCreateWithFlags(&ctx->ev_run[i], cudaEventDisableTiming));
            }
        }
        // a chunk is a whole number of waves (one CTA per stream, num_sms * occ resident): anything else pays
        // for a second, nearly empty wave; the last chunk takes the remainder
        const int wave = ctx->num_sms * occ;
        int per = wave * ((nstreams + 16 * wave - 1) / (16 * wave));
        int nchunks = nstreams / per;
        if (nchunks < 1) nchunks = 1;
        AC3_CUDA(cudaMemsetAsync(ctx->d_counter, 0, 16 * sizeof(int), ctx->s_run));
        if (carry) AC3_CUDA(cudaMemcpyAsync(ctx->b_carry.p, carry, sizeof(EncCarry) * (size_t)nstreams, cudaMemcpyHostToDevice, ctx->s_run));
        const size_t pcm_per = (size_t)nframes * 1536 * channels * 2, out_per = (size_t)nframes * c.frame_words * 2;
        for (int k = 0; k < nchunks; k++) {
            const int s0 = k * per, s1 = (k + 1 == nchunks) ? nstreams : (k + 1) * per;
            AC3_CUDA(cudaMemcpyAsync((uint8_t*)ctx->b_pcm.p + s0 * pcm_per, (const uint8_t*)pcm + s0 * pcm_per,
                                     (s1 - s0) * pcm_per, cudaMemcpyHostToDevice, ctx->s_in));
            AC3_CUDA(cudaEventRecord(ctx->ev_in[k], ctx->s_in));
            AC3_CUDA(cudaStreamWaitEvent(ctx->s_run, ctx->ev_in[k], 0));
            EncParams Pk = P;
            Pk.pcm = P.pcm + (size_t)s0 * nframes * 1536 * channels;
            Pk.out = P.out + s0 * out_per;
            Pk.status = P.status + (size_t)s0 * nframes;
            Pk.carry = P.carry ? P.carry + s0 : nullptr;
            Pk.slice_done = P.slice_done ? P.slice_done + s0 : nullptr;
            Pk.nstreams = s1 - s0;
            Pk.work_counter = ctx->d_counter + k;
            int gk = ctx->num_sms * occ;
            if (gk > s1 - s0) gk = s1 - s0;
            ac3_encode_kernel<<<gk, kThreads, smem, ctx->s_run>>>(Pk);
            AC3_CUDA(cudaGetLastError());
            ctx->launches++;
            AC3_CUDA(cudaEventRecord(ctx->ev_run[k], ctx->s_run));
            AC3_CUDA(cudaStreamWaitEvent(ctx->s_out, ctx->ev_run[k], 0));
            AC3_CUDA(cudaMemcpyAsync(out + s0 * out_per, P.out + s0 * out_per, (s1 - s0) * out_per, cudaMemcpyDeviceToHost, ctx->s_out));
            if (status)
                AC3_CUDA(cudaMemcpyAsync(status + (size_t)s0 * nframes, P.status + (size_t)s0 * nframes,
                                         (size_t)(s1 - s0) * nframes * 4, cudaMemcpyDeviceToHost, ctx->s_out));
        }
        if (carry) AC3_CUDA(cudaMemcpyAsync(carry, ctx->b_carry.p, sizeof(EncCarry) * (size_t)nstreams, cudaMemcpyDeviceToHost, ctx->s_run));
        AC3_CUDA(cudaStreamSynchronize(ctx->s_run));
        AC3_CUDA(cudaStreamSynchronize(ctx->s_out));
        AC3_CUDA(cudaStreamSynchronize(ctx->s_in));
        return 0;
    }
    AC3_CUDA(cudaMemsetAsync(ctx->d_counter, 0, sizeof(int), st));
    if (ctx->ev_used + 2 > ctx->ev_pool.size()) {
        if (ctx->ev_pool.size() < 8192) {
            cudaEvent_t a, b;
            AC3_CUDA(cudaEventCreate(&a));
            AC3_CUDA(cudaEventCreate(&b));
            ctx->ev_pool.push_back(a);
            ctx->ev_pool.push_back(b);
        } else ctx->ev_used = 0;
    }
    cudaEvent_t e0 = ctx->ev_pool[ctx->ev_used], e1 = ctx->ev_pool[ctx->ev_used + 1];
    ctx->ev_used += 2;
    AC3_CUDA(cudaEventRecord(e0, st));
    ac3_encode_kernel<<<grid, kThreads, smem, st>>>(P);
    AC3_CUDA(cudaEventRecord(e1, st));
    AC3_CUDA(cudaGetLastError());
    ctx->launches++;
    if (dev) return 0;

    AC3_CUDA(cudaMemcpyAsync(out, P.out, out_bytes, cudaMemcpyDeviceToHost, st));
    if (status) AC3_CUDA(cudaMemcpyAsync(status, P.status, total * 4, cudaMemcpyDeviceToHost, st));
    if (carry) AC3_CUDA(cudaMemcpyAsync(carry, ctx->b_carry.p, sizeof(EncCarry) * (size_t)nstreams, cudaMemcpyDeviceToHost, st));
    if (P.dbg_coef && debug) {
        AC3_CUDA(cudaMemcpyAsync(debug->coef, P.dbg_coef, total * 9216 * 4, cudaMemcpyDeviceToHost, st));
        AC3_CUDA(cudaMemcpyAsync(debug->exp_shift, P.dbg_shift, total * 36, cudaMemcpyDeviceToHost, st));
        AC3_CUDA(cudaMemcpyAsync(debug->strategy, P.dbg_strategy, total * 36, cudaMemcpyDeviceToHost, st));
        AC3_CUDA(cudaMemcpyAsync(debug->encoded_exp, P.dbg_enc, total * 9216, cudaMemcpyDeviceToHost, st));
        AC3_CUDA(cudaMemcpyAsync(debug->bap, P.dbg_bap, total * 9216, cudaMemcpyDeviceToHost, st));
        AC3_CUDA(cudaMemcpyAsync(debug->snr, P.dbg_snr, total * 8, cudaMemcpyDeviceToHost, st));
    }
    AC3_CUDA(cudaStreamSynchronize(st));
    return 0;
}

// ===========================================================================
// encoder front end: the part of the ACM wrapper that sits around AC3_encode_frame
// ===========================================================================
int ac3_wav_channel_map(int channels, uint8_t chmap[6])      // AC3ACM.cpp:1631-1662
{
    // WAVE order FL FR FC LFE BL BR; coded order L C R SL SR LFE
    static const uint8_t with_centre[6] = {0, 2, 1, 3, 4, 0}, six[6] = {0, 2, 1, 4, 5, 3};
    switch (channels) {
    case 1: case 2: case 4:
        for (int i = 0; i < 4; i++) chmap[i] = (uint8_t)i;
        return 0;
    case 3: case 5:
        for (int i = 0; i < 5; i++) chmap[i] = with_centre[i];
        return 0;
    case 6:
        for (int i = 0; i < 6; i++) chmap[i] = six[i];
        return 0;
    }
    return -1;
}

// 16-bit words per frame at 32 / 44.1 / 48 kHz for the 19 bitrates (AC3ACM.cpp:128-149); 44.1 kHz: the
// unpadded size, which is what the encoder emits (ac3enc.cpp:1076-1077)
static const uint16_t acm_kbps[19] = {32, 40, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320, 384, 448, 512, 576, 640};

static int acm_frame_words(int rate_index, int i)            // rate_index: 0 = 32 k, 1 = 44.1 k, 2 = 48 k
{
    const int kbps = acm_kbps[i];
    return rate_index == 0 ? 3 * kbps : rate_index == 2 ? 2 * kbps : 320 * kbps / 147;
}

int ac3_acm_bitrate(int freq, uint32_t avg_bytes_per_sec)    // AC3ACM.cpp:1913-1936
{
    const uint32_t fl = avg_bytes_per_sec / 125;
    for (int i = 0; i < 19; i++)
        if (fl == acm_kbps[i]) return acm_kbps[i];
    if (freq == 44100)
        for (int i = 0; i < 19; i++)
            if ((uint32_t)((acm_frame_words(1, i) * 2 * 44100 + 768) / 1536) == avg_bytes_per_sec) return acm_kbps[i];
    return 0;
}

int ac3_acm_block_align(int freq, int kbps)                  // AC3ACM.cpp:951-953
{
    const int ri = freq == 32000 ? 0 : freq == 44100 ? 1 : freq == 48000 ? 2 : -1;
    if (ri < 0) return 0;
    for (int i = 0; i < 19; i++)
        if (acm_kbps[i] == kbps) return acm_frame_words(ri, i) * 2;
    return 0;
}

struct ac3_stream_s {
    ac3_batch_t* ctx;
    int freq, bitrate, channels, frame_bytes, needed;
    uint8_t chmap[8];
    ac3_stream_carry_t carry;
    std::vector<uint8_t> buf;        // msd->buf: input gathered so far (< needed bytes between calls)
    std::vector<uint8_t> frame;      // msd->frame: the last frame encoded
    int fill;                        // msd->bufptr - msd->buf
    int pending, pending_at;         // msd->blocks bytes of `frame` still owed, from msd->bufend
    std::vector<int16_t> stage;
    std::vector<uint8_t> out;
};

ac3_stream_t* ac3_stream_open(ac3_batch_t* ctx, int freq, uint32_t avg_bytes_per_sec, int channels)
{
    if (!ctx || freq < 32000) return nullptr;                // AC3ACM.cpp:1890-1891
    const int kbps = ac3_acm_bitrate(freq, avg_bytes_per_sec);
    if (!kbps) return nullptr;
    const int fb = ac3_batch_frame_bytes(freq, kbps * 1000, channels);
    if (!fb) return nullptr;
    ac3_stream_t* s = new ac3_stream_s();
    s->ctx = ctx;
    s->freq = freq;
    s->bitrate = kbps * 1000;
    s->channels = channels;
    s->frame_bytes = fb;
    s->needed = 1536 * channels * (int)sizeof(int16_t);
    memset(s->chmap, 0, sizeof(s->chmap));
    ac3_wav_channel_map(channels, s->chmap);
    memset(&s->carry, 0, sizeof(s->carry));
    s->buf.resize(s->needed);
    s->frame.resize(fb + 8);
    s->fill = s->pending = s->pending_at = 0;
    return s;
}

int ac3_stream_frame_bytes(ac3_stream_t* s) { return s ? s->frame_bytes : 0; }

void ac3_stream_close(ac3_stream_t* s) { delete s; }

int ac3_stream_convert(ac3_stream_t* s, const void* src_, uint32_t src_len, uint32_t* src_used, void* dst_,
                       uint32_t dst_len, uint32_t* dst_used, int start)
{
    if (!s) return -1;
    const uint8_t* src = (const uint8_t*)src_;
    uint8_t* dst = (uint8_t*)dst_;
    uint32_t sused = 0, dused = 0;
    long srcLen = src_len, dstLen = dst_len;
    if (start) {                                             // ACM_STREAMCONVERTF_START (:1705-1710)
        s->fill = 0;
        s->pending = s->pending_at = 0;
    } else if (s->pending > 0) {                             // the rest of the last frame first (:1711-1734)
        long tc = s->pending < dstLen ? s->pending : dstLen;
        if (tc > 0) {
            memcpy(dst, s->frame.data() + s->pending_at, tc);
            dused += tc;
            s->pending -= tc;
            s->pending_at += tc;
            dstLen -= tc;
            dst += tc;
        }
        if (dstLen <= 0) {
            if (src_used) *src_used = 0;
            if (dst_used) *dst_used = dused;
            return 0;
        }
    }
    // Walk the reference's loop (:1743-1785) without encoding: it takes input until the gather buffer is full,
    // encodes, hands out what fits, and stops after the frame that exhausts the destination.
    int nframes = 0, fill = s->fill;
    long consumed = 0;
    {
        long remaining = srcLen, room = dstLen;
        while (remaining > 0) {
            long tc = s->needed - fill;
            if (tc > remaining) tc = remaining;
            remaining -= tc;
            consumed += tc;
            fill += (int)tc;
            if (fill >= s->needed) {
                nframes++;
                fill = 0;
                room -= (s->frame_bytes < room ? s->frame_bytes : room);
                if (room <= 0) break;
            }
        }
    }
    long tail_from = 0;                                      // input offset of the bytes that stay buffered
    if (nframes > 0) {
        // all frames of this call in one launch: the buffered head + input bytes
        s->stage.resize((size_t)nframes * 1536 * s->channels);
        uint8_t* st = (uint8_t*)s->stage.data();
        memcpy(st, s->buf.data(), s->fill);
        const long take = (long)nframes * s->needed - s->fill;
        memcpy(st + s->fill, src, take);
        tail_from = take;
        s->fill = 0;
        s->out.resize((size_t)nframes * s->frame_bytes);
        int rc = ac3_batch_encode(s->ctx, s->stage.data(), 1, nframes, s->freq, s->bitrate, s->channels, s->chmap,
                                  s->out.data(), nullptr, &s->carry, nullptr, 0, nullptr);
        if (rc) return rc;
        for (int f = 0; f < nframes; f++) {
            const uint8_t* fr = s->out.data() + (size_t)f * s->frame_bytes;
            long tc = s->frame_bytes < dstLen ? s->frame_bytes : dstLen;
            if (tc > 0) memcpy(dst, fr, tc);
            dused += tc;
            dst += tc;
            dstLen -= tc;
            if (f == nframes - 1) {                          // msd->frame / msd->blocks / msd->bufend
                memcpy(s->frame.data(), fr, s->frame_bytes);
                s->pending = s->frame_bytes - (int)tc;
                s->pending_at = (int)tc;
            }
        }
    }
    memcpy(s->buf.data() + s->fill, src + tail_from, consumed - tail_from);
    s->fill += (int)(consumed - tail_from);
    sused += consumed;
    if (src_used) *src_used = sused;
    if (dst_used) *dst_used = dused;
    return 0;
}

// ===========================================================================
// drop-in encoder API (include/ac3enc.h): a process-wide singleton like the reference's
// (static AC3EncodeContext ac3enc_state, ac3enc.cpp:78)
// ===========================================================================
static struct {
    ac3_batch_t* ctx;
    int freq, bitrate, channels, frame_bytes;
    ac3_stream_carry_t carry;
} g_enc1;

int AC3_encode_init(int freq, int bitrate, int channels)
{
    int fb = ac3_batch_frame_bytes(freq, bitrate, channels);
    if (!fb) return 0;
    if (!g_enc1.ctx) {
        int dev = 0;
        const char* e = getenv("A52_B200_DEVICE");
        if (e) dev = atoi(e);
        g_enc1.ctx = ac3_batch_create(dev);
        if (!g_enc1.ctx) return 0;          // no CPU fallback by design
    }
    g_enc1.freq = freq;
    g_enc1.bitrate = bitrate;
    g_enc1.channels = channels;
    g_enc1.frame_bytes = fb;
    // the reference keeps last_samples across re-initialisation but resets the snr warm start (:1092);
    // a fresh stream here starts from silence like a freshly loaded codec
    memset(&g_enc1.carry, 0, sizeof(g_enc1.carry));
    return fb;
}

int AC3_encode_frame(unsigned char* dst, short* samples, unsigned char* chmap)
{
    if (!g_enc1.ctx) return 0;
    int rc = ac3_batch_encode(g_enc1.ctx, samples, 1, 1, g_enc1.freq, g_enc1.bitrate, g_enc1.channels, chmap, dst,
                              nullptr, &g_enc1.carry, nullptr, 0, nullptr);
    return rc ? 0 : g_enc1.frame_bytes;
}

}  // extern "C"
#pragma GCC visibility pop
