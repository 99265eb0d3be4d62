// Per-launch CUDA-event timing for bench.py (include/rtda_lovasz.h, "Per-kernel timing").
#include <vector>

#include "common.cuh"

namespace rtda {

struct Span { int stage; cudaEvent_t a, b; };
static bool g_on = false;
static std::vector<Span> g_spans;
static std::vector<cudaEvent_t> g_pool;
static cudaEvent_t g_open[RTDA_PROFILE_STAGES];

static cudaEvent_t get_event() {
    if (!g_pool.empty()) { cudaEvent_t e = g_pool.back(); g_pool.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

bool profile_enabled() { return g_on; }

void profile_begin(int stage, cudaStream_t st) {
    cudaEvent_t a = get_event();
    cudaEventRecord(a, st);
    g_open[stage] = a;
}

void profile_end(int stage, cudaStream_t st) {
    cudaEvent_t b = get_event();
    cudaEventRecord(b, st);
    g_spans.push_back(Span{stage, g_open[stage], b});
}

}  // namespace rtda

using namespace rtda;

extern "C" int rtda_profile_enable(int on) {
    g_on = on != 0;
    return 0;
}

extern "C" int rtda_profile_read(double *ms, int64_t *launches, int n_stages) {
    if (!ms || !launches || n_stages < RTDA_PROFILE_STAGES) return RTDA_E_ARG;
    RTDA_RETURN_IF_CUDA_ERROR(cudaDeviceSynchronize());
    const Span *prev = nullptr;
    for (const Span &s : g_spans) {
        float t = 0.f;
        RTDA_RETURN_IF_CUDA_ERROR(cudaEventElapsedTime(&t, s.a, s.b));
        ms[s.stage] += (double)t;
        launches[s.stage] += 1;
        if (prev) {  // time between two consecutive launches of this library (idle, or someone else's kernels)
            float gap = 0.f;
            if (cudaEventElapsedTime(&gap, prev->b, s.a) == cudaSuccess && gap > 0.f) {
                ms[kStageOther] += (double)gap;
                launches[kStageOther] += 1;
            }
        }
        prev = &s;
    }
    for (const Span &s : g_spans) {
        g_pool.push_back(s.a);
        g_pool.push_back(s.b);
    }
    g_spans.clear();
    return 0;
}

extern "C" const char *rtda_profile_stage_name(int stage) {
    static const char *names[RTDA_PROFILE_STAGES] = {
        "memset_workspace", "k_prep", "k_scan", "k_sort_pass<0,probas>", "k_sort_pass<1>", "k_sort_pass<2>",
        "k_rank_last", "k_finalize", "k_bwd", "k_confusion", "k_confusion_argmax", "gaps_between_launches",
        "k_split", "k_part", "k_local", "k_hist_fb", "k_plan"};
    if (stage < 0 || stage >= RTDA_PROFILE_STAGES) return "?";
    return names[stage];
}
