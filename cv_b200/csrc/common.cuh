// cv_b200/csrc/common.cuh -- shared declarations for libcvb200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include "../../include/cvb200.h"

struct AkazeWorkspace;
struct MatchWorkspace;
struct GeomWorkspace;
struct PairWorkspace;

struct cvb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    std::string err;
    uint64_t launches = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t ev_wait = nullptr;   // blocking-sync event of cvb_wait
    int num_sms = 148;
    AkazeWorkspace *akaze = nullptr;
    MatchWorkspace *match = nullptr;
    GeomWorkspace *geom = nullptr;
    PairWorkspace *pair = nullptr;
    // page-locked host scratch for the small device->host results of the host API (a D2H copy into pageable memory is
    // staged synchronously inside the driver and stalls the other contexts' launches)
    void *pinned = nullptr;
    size_t pinned_bytes = 0;
    // optional per-kernel CUDA-event profiling (bench.py roofline pass); off by default
    bool prof = false;
    std::vector<cudaEvent_t> prof_pool;
    size_t prof_used = 0;
    struct ProfRec { const char *name; cudaEvent_t e0, e1; double bytes; };
    std::vector<ProfRec> prof_recs;
};

cudaEvent_t cvb_prof_event(cvb_ctx *ctx);
// Host wait for a stream.  Default: the thread sleeps on a blocking-sync event (a service runs one host thread per context and
// several ranks per box; spinning threads take the cores the launching threads need).  CVB_SYNC=spin: cudaStreamSynchronize.
cudaError_t cvb_wait(cvb_ctx *ctx, cudaStream_t st);
void *cvb_pinned(cvb_ctx *ctx, size_t bytes);   // >= bytes of page-locked scratch (nullptr on failure); valid until the next call
struct CvbProfScope {
    cvb_ctx *ctx; const char *name; double bytes; cudaEvent_t e0 = nullptr;
    CvbProfScope(cvb_ctx *c, const char *n, double b) : ctx(c), name(n), bytes(b) {
        if (ctx->prof) { e0 = cvb_prof_event(ctx); cudaEventRecord(e0, ctx->stream); }
    }
    ~CvbProfScope() {
        if (ctx->prof && e0) { cudaEvent_t e1 = cvb_prof_event(ctx); cudaEventRecord(e1, ctx->stream); ctx->prof_recs.push_back({name, e0, e1, bytes}); }
    }
};
// PROF(ctx, "kernel", algorithmic_bytes) brackets the launches that follow in the current scope
#define CVB_PROF(ctx, name, bytes) CvbProfScope prof_scope__((ctx), (name), (double)(bytes))

int cvb_set_error(cvb_ctx *ctx, int code, const char *fmt, ...);
void akaze_workspace_free(AkazeWorkspace *ws);
void match_workspace_free(MatchWorkspace *ws);
void geom_workspace_free(GeomWorkspace *ws);
void pair_workspace_free(PairWorkspace *ws);

#define CVB_CUDA(ctx, call)                                                                          \
    do {                                                                                             \
        cudaError_t e__ = (call);                                                                    \
        if (e__ != cudaSuccess)                                                                      \
            return cvb_set_error((ctx), CVB_ECUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call,        \
                                 cudaGetErrorString(e__));                                           \
    } while (0)

#define CVB_LAUNCH_CHECK(ctx)                                                                        \
    do {                                                                                             \
        (ctx)->launches++;                                                                           \
        cudaError_t e__ = cudaPeekAtLastError();                                                     \
        if (e__ != cudaSuccess)                                                                      \
            return cvb_set_error((ctx), CVB_ECUDA, "%s:%d launch: %s", __FILE__, __LINE__,           \
                                 cudaGetErrorString(e__));                                           \
    } while (0)

static inline unsigned cdiv(unsigned a, unsigned b) { return (a + b - 1) / b; }
