// rtk_alloc.cpp — operator new / delete of librtk_b200.so: a small thread-caching allocator.
//
// The region logic (correct.cpp / traverse.cpp) copies and spells millions of 1-2 kB strings and path vectors per second
// from ~23 threads; with glibc's allocator a third of the host time of a correction step was malloc / free (arena locks:
// blocks are typically allocated by a worker and freed by a service thread or vice versa).  Here every thread keeps
// free lists per size class and touches a lock only to exchange 32 blocks at a time with a global pool.  Blocks carry a
// 16-byte header (class); anything above 8 KiB goes to malloc.  Memory handed to a size class is never returned to the
// system (bounded by the peak working set of a batch).  The library is linked with -static-libstdc++ so that EVERY
// allocation of its C++ code (including libstdc++'s out-of-line string code) pairs with these functions; no C++ object
// crosses the C ABI (outputs are malloc'ed, include/rtk.h).  RTK_SYSTEM_MALLOC=1 falls back to malloc for new blocks.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>

namespace {

constexpr int kClasses = 20;
constexpr size_t kSizes[kClasses] = {16, 32, 48, 64, 96, 128, 192, 256, 384, 512, 768, 1024, 1536, 2048, 3072, 4096, 5120, 6144, 7168, 8192};
constexpr size_t kMaxSmall = 8192;
constexpr size_t kHeader = 16;
constexpr uint32_t kLarge = 0xFFFFFFFFu;
constexpr size_t kBatch = 32;          // blocks moved between a thread cache and the global pool at once
constexpr size_t kCacheMax = 4 * kBatch;

struct Header { uint32_t cls; uint32_t magic; uint64_t pad; };
struct FreeNode { FreeNode* next; };

inline int class_of(size_t n) {
    // n <= kMaxSmall
    if (n <= 64) return (int)((n + 15) / 16) - 1;                 // 16, 32, 48, 64
    int c = 4;
    while (kSizes[c] < n) ++c;
    return c;
}

struct GlobalPool {
    std::mutex mu;
    FreeNode* head = nullptr;   // singly linked list of free blocks (payload pointers)
};
GlobalPool g_pool[kClasses];
bool g_system = false;

struct ThreadCache {
    FreeNode* head[kClasses] = {nullptr};
    uint32_t count[kClasses] = {0};
    ~ThreadCache() {   // thread exit: everything back to the global pool
        for (int c = 0; c < kClasses; ++c) {
            if (!head[c]) continue;
            FreeNode* last = head[c];
            while (last->next) last = last->next;
            std::lock_guard<std::mutex> lk(g_pool[c].mu);
            last->next = g_pool[c].head;
            g_pool[c].head = head[c];
            head[c] = nullptr; count[c] = 0;
        }
    }
};
thread_local ThreadCache tl_cache;

// carve a slab of `kBatch` blocks of class c for the calling thread
inline void refill(ThreadCache& tc, int c) {
    {   // first the global pool
        std::lock_guard<std::mutex> lk(g_pool[c].mu);
        size_t got = 0;
        while (g_pool[c].head && got < kBatch) {
            FreeNode* n = g_pool[c].head;
            g_pool[c].head = n->next;
            n->next = tc.head[c];
            tc.head[c] = n;
            ++got;
        }
        tc.count[c] += (uint32_t)got;
        if (got) return;
    }
    const size_t bs = kHeader + kSizes[c];
    const size_t nblk = kSizes[c] <= 1024 ? 64 : 16;
    char* slab = (char*)malloc(bs * nblk);
    if (!slab) throw std::bad_alloc();
    for (size_t i = 0; i < nblk; ++i) {
        Header* h = (Header*)(slab + i * bs);
        h->cls = (uint32_t)c; h->magic = 0x52544B41u; h->pad = 0;
        FreeNode* n = (FreeNode*)(slab + i * bs + kHeader);
        n->next = tc.head[c];
        tc.head[c] = n;
    }
    tc.count[c] += (uint32_t)nblk;
}

inline void* rtk_new(size_t n) {
    if (n == 0) n = 1;
    if (n > kMaxSmall || g_system) {
        char* p = (char*)malloc(n + kHeader);
        if (!p) throw std::bad_alloc();
        Header* h = (Header*)p;
        h->cls = kLarge; h->magic = 0x52544B41u; h->pad = 0;
        return p + kHeader;
    }
    const int c = class_of(n);
    ThreadCache& tc = tl_cache;
    if (!tc.head[c]) refill(tc, c);
    FreeNode* b = tc.head[c];
    tc.head[c] = b->next;
    --tc.count[c];
    return (void*)b;
}

inline void rtk_delete(void* p) noexcept {
    if (!p) return;
    Header* h = (Header*)((char*)p - kHeader);
    if (h->cls == kLarge) { free(h); return; }
    const int c = (int)h->cls;
    ThreadCache& tc = tl_cache;
    FreeNode* n = (FreeNode*)p;
    n->next = tc.head[c];
    tc.head[c] = n;
    if (++tc.count[c] > kCacheMax) {   // give half back
        FreeNode* first = tc.head[c];
        FreeNode* last = first;
        for (size_t i = 1; i < kCacheMax / 2; ++i) last = last->next;
        tc.head[c] = last->next;
        tc.count[c] -= (uint32_t)(kCacheMax / 2);
        std::lock_guard<std::mutex> lk(g_pool[c].mu);
        last->next = g_pool[c].head;
        g_pool[c].head = first;
    }
}

struct Init { Init() { g_system = getenv("RTK_SYSTEM_MALLOC") != nullptr; } } g_init;

}  // namespace

void* operator new(size_t n) { return rtk_new(n); }
void* operator new[](size_t n) { return rtk_new(n); }
void* operator new(size_t n, const std::nothrow_t&) noexcept { try { return rtk_new(n); } catch (...) { return nullptr; } }
void* operator new[](size_t n, const std::nothrow_t&) noexcept { try { return rtk_new(n); } catch (...) { return nullptr; } }
void operator delete(void* p) noexcept { rtk_delete(p); }
void operator delete[](void* p) noexcept { rtk_delete(p); }
void operator delete(void* p, size_t) noexcept { rtk_delete(p); }
void operator delete[](void* p, size_t) noexcept { rtk_delete(p); }
void operator delete(void* p, const std::nothrow_t&) noexcept { rtk_delete(p); }
void operator delete[](void* p, const std::nothrow_t&) noexcept { rtk_delete(p); }
// over-aligned types are not used by the library; keep the pairs consistent anyway
void* operator new(size_t n, std::align_val_t a) { void* p = aligned_alloc((size_t)a, (n + (size_t)a - 1) / (size_t)a * (size_t)a); if (!p) throw std::bad_alloc(); return p; }
void* operator new[](size_t n, std::align_val_t a) { return operator new(n, a); }
void operator delete(void* p, std::align_val_t) noexcept { free(p); }
void operator delete[](void* p, std::align_val_t) noexcept { free(p); }
void operator delete(void* p, size_t, std::align_val_t) noexcept { free(p); }
void operator delete[](void* p, size_t, std::align_val_t) noexcept { free(p); }
