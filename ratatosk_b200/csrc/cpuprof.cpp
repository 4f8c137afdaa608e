// cpuprof.cpp — development aid: in-process sampling CPU profiler (SIGPROF + backtrace), enabled by
// RTK_CPU_PROFILE=<file>.  Samples every thread of the process in proportion to its CPU time; the raw return
// addresses and /proc/self/maps are written at exit and symbolised offline (scripts/cpuprof_report.py).
#include <execinfo.h>
#include <signal.h>
#include <sys/time.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace {
constexpr int kDepth = 20;
constexpr size_t kMax = 1u << 20;
void* (*g_samples)[kDepth] = nullptr;
std::atomic<size_t> g_n{0};
const char* g_path = nullptr;

void on_prof(int, siginfo_t*, void*) {
    const size_t i = g_n.fetch_add(1);
    if (i >= kMax) return;
    void* tmp[kDepth + 2];
    const int n = backtrace(tmp, kDepth + 2);
    for (int j = 0; j < kDepth; ++j) g_samples[i][j] = (j + 2 < n) ? tmp[j + 2] : nullptr;   // skip the handler frames
}

void dump() {
    struct itimerval off;
    memset(&off, 0, sizeof(off));
    setitimer(ITIMER_PROF, &off, nullptr);
    FILE* f = fopen(g_path, "w");
    if (!f) return;
    FILE* m = fopen("/proc/self/maps", "r");
    char line[1024];
    while (m && fgets(line, sizeof(line), m)) if (strstr(line, " r-xp ")) fprintf(f, "M %s", line);
    if (m) fclose(m);
    const size_t n = g_n.load() < kMax ? g_n.load() : kMax;
    for (size_t i = 0; i < n; ++i) {
        fputc('S', f);
        for (int j = 0; j < kDepth && g_samples[i][j]; ++j) fprintf(f, " %p", g_samples[i][j]);
        fputc('\n', f);
    }
    fclose(f);
}

struct Init {
    Init() {
        g_path = getenv("RTK_CPU_PROFILE");
        if (!g_path || !*g_path) return;
        g_samples = (void* (*)[kDepth])calloc(kMax, sizeof(void*) * kDepth);
        if (!g_samples) return;
        void* warm[4];
        backtrace(warm, 4);   // loads the unwinder outside signal context
        struct sigaction sa;
        memset(&sa, 0, sizeof(sa));
        sa.sa_sigaction = on_prof;
        sa.sa_flags = SA_RESTART | SA_SIGINFO;
        sigaction(SIGPROF, &sa, nullptr);
        struct itimerval it;
        it.it_interval.tv_sec = 0; it.it_interval.tv_usec = 997;
        it.it_value = it.it_interval;
        setitimer(ITIMER_PROF, &it, nullptr);
        atexit(dump);
    }
} g_init;
}  // namespace
