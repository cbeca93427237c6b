// pb_ctx.cu -- context, error reporting, stage timers, device buffers.
#include <stdarg.h>
#include <stdlib.h>

#include "pb_internal.cuh"

static thread_local std::string g_create_error;

void pb_set_error(pb_ctx *ctx, const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf; else g_create_error = buf;
}

int pb_fail(pb_ctx *ctx, int code, const char *fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf; else g_create_error = buf;
    return code;
}

extern "C" int pb_abi_version(void) { return PB_ABI_VERSION; }

extern "C" int pb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

extern "C" int pb_ctx_create(int device, pb_ctx **out)
{
    if (!out) return pb_fail(nullptr, PB_ERR_ARG, "pb_ctx_create: out is NULL");
    *out = nullptr;
    // The aligner runs one kernel per band class on its own stream (up to 14); the default of 8 hardware queues would
    // serialise the rest.  Only effective if CUDA has not been initialised in this process yet (bench.py sets it too).
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
    int n = pb_device_count();
    if (n <= 0)
        return pb_fail(nullptr, PB_ERR_NO_DEVICE,
                       "no CUDA device visible: this library has no CPU fallback (sm_100a kernels only)");
    if (device < 0 || device >= n) return pb_fail(nullptr, PB_ERR_ARG, "device %d out of range [0,%d)", device, n);
    pb_ctx *ctx = new pb_ctx();
    ctx->device = device;
    PB_CUDA(nullptr, cudaSetDevice(device));
    cudaDeviceProp prop;
    PB_CUDA(nullptr, cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        delete ctx;
        return pb_fail(nullptr, PB_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", device,
                       prop.major, prop.minor);
    }
    ctx->sm_count = prop.multiProcessorCount;
    // The traceback reads 8-byte parent pairs scattered over rows: keep DRAM->L2 fills at sector size (a hint)
    {
        const char *g = getenv("PB_L2_FETCH");
        cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, g ? (size_t)atoi(g) : 32);
        cudaGetLastError();
    }
    PB_CUDA(nullptr, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2 * PB_T_COUNT; ++i) PB_CUDA(nullptr, cudaEventCreate(&ctx->ev[i]));
    // keep freed blocks in the pool: per-step allocations become pointer bumps
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t thr = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    ctx->h_pin_bytes = 1 << 20;
    PB_CUDA(nullptr, cudaMallocHost(&ctx->h_pin, ctx->h_pin_bytes));
    *out = ctx;
    return PB_OK;
}

extern "C" void pb_ctx_destroy(pb_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (int i = 0; i < 2 * PB_T_COUNT; ++i)
        if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    if (ctx->h_pin) cudaFreeHost(ctx->h_pin);
    for (auto st : ctx->aux_streams) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
    for (auto ev : ctx->aux_events) cudaEventDestroy(ev);
    if (ctx->fork_event) cudaEventDestroy(ctx->fork_event);
    if (ctx->scratch) cudaFree(ctx->scratch);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" const char *pb_last_error(const pb_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }
extern "C" void *pb_ctx_stream(pb_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
extern "C" int64_t pb_ctx_launch_count(const pb_ctx *ctx) { return ctx ? ctx->launches : 0; }

extern "C" int pb_ctx_set_scratch_limit(pb_ctx *ctx, size_t bytes)
{
    if (!ctx) return PB_ERR_ARG;
    ctx->scratch_limit = bytes;
    return PB_OK;
}

extern "C" int pb_ctx_timings(const pb_ctx *ctx, float *ms)
{
    if (!ctx || !ms) return PB_ERR_ARG;
    for (int i = 0; i < PB_T_COUNT; ++i) ms[i] = ctx->times[i];
    return PB_OK;
}

void pb_timer_reset(pb_ctx *ctx)
{
    for (int i = 0; i < PB_T_COUNT; ++i) { ctx->times[i] = 0.f; ctx->timed[i] = false; }
}
void pb_timer_begin(pb_ctx *ctx, int which) { cudaEventRecord(ctx->ev[2 * which], ctx->stream); }
void pb_timer_end(pb_ctx *ctx, int which)
{
    cudaEventRecord(ctx->ev[2 * which + 1], ctx->stream);
    ctx->timed[which] = true;
}
void pb_timer_collect(pb_ctx *ctx)
{
    for (int i = 0; i < PB_T_COUNT; ++i)
        if (ctx->timed[i]) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, ctx->ev[2 * i], ctx->ev[2 * i + 1]) == cudaSuccess) ctx->times[i] = ms;
            else cudaGetLastError();
        }
}

int DevBuf::alloc(pb_ctx *c, size_t n)
{
    release();
    ctx = c;
    if (n == 0) n = 16;
    cudaError_t e = cudaMallocAsync(&p, n, c->stream);
    if (e != cudaSuccess) {
        p = nullptr;
        cudaGetLastError();
        return pb_fail(c, PB_ERR_NOMEM, "device allocation of %zu bytes failed: %s", n, cudaGetErrorString(e));
    }
    bytes = n;
    return PB_OK;
}

int DevBuf::alloc_zero(pb_ctx *c, size_t n)
{
    PB_TRY(alloc(c, n));
    PB_CUDA(c, cudaMemsetAsync(p, 0, bytes, c->stream));
    return PB_OK;
}

void DevBuf::release()
{
    if (p && ctx) cudaFreeAsync(p, ctx->stream);
    p = nullptr;
    bytes = 0;
}

int pb_h2d(pb_ctx *ctx, void *dst, const void *src, size_t bytes)
{
    if (!bytes) return PB_OK;
    PB_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return PB_OK;
}

int pb_d2h(pb_ctx *ctx, void *dst, const void *src, size_t bytes)
{
    if (!bytes) return PB_OK;
    PB_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return PB_OK;
}

int pb_sync(pb_ctx *ctx)
{
    PB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PB_OK;
}
