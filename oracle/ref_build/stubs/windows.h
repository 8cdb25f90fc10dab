/*
 * windows.h stub -- TEST INFRASTRUCTURE.  The handful of Win32 names the reference's work queue
 * uses (work_queue.cpp:15-106, demofox_path_tracing_optimization_v4.cpp:1567-1694), mapped to
 * POSIX: semaphore -> sem_t, CreateThread -> detached pthread, Interlocked* -> __sync builtins.
 */
#ifndef ORACLE_STUB_WINDOWS_H
#define ORACLE_STUB_WINDOWS_H
#include <pthread.h>
#include <semaphore.h>
#include <stdint.h>
#include <stdlib.h>

typedef void* HANDLE;
typedef uint32_t DWORD;
typedef int32_t LONG;
typedef void* LPVOID;
typedef const char* LPCSTR;
typedef void* LPSECURITY_ATTRIBUTES;
#define WINAPI
#define INFINITE 0xFFFFFFFFu
#ifndef FALSE
#define FALSE 0
#endif
#define SEMAPHORE_ALL_ACCESS 0x1F0003

static inline HANDLE CreateSemaphoreExA(LPSECURITY_ATTRIBUTES, LONG initial, LONG, LPCSTR, DWORD, DWORD)
{
    sem_t* s = (sem_t*)malloc(sizeof(sem_t));
    sem_init(s, 0, (unsigned)initial);
    return (HANDLE)s;
}
static inline int ReleaseSemaphore(HANDLE h, LONG count, LONG* prev)
{
    if (prev) *prev = 0;
    for (LONG i = 0; i < count; i++) sem_post((sem_t*)h);
    return 1;
}
static inline DWORD WaitForSingleObjectEx(HANDLE h, DWORD, int)
{
    sem_wait((sem_t*)h);
    return 0;
}
struct oracle_thread_start { DWORD (*proc)(LPVOID); LPVOID param; };
static inline void* oracle_thread_trampoline(void* p)
{
    oracle_thread_start s = *(oracle_thread_start*)p;
    free(p);
    s.proc(s.param);
    return 0;
}
static inline HANDLE CreateThread(void*, size_t, DWORD (*proc)(LPVOID), LPVOID param, DWORD, DWORD* id)
{
    oracle_thread_start* s = (oracle_thread_start*)malloc(sizeof(oracle_thread_start));
    s->proc = proc;
    s->param = param;
    pthread_t t;
    if (pthread_create(&t, 0, oracle_thread_trampoline, s) != 0) return 0;
    pthread_detach(t);
    if (id) *id = 0;
    return (HANDLE)1;
}
static inline int CloseHandle(HANDLE) { return 1; }
static inline LONG InterlockedCompareExchange(LONG volatile* dst, LONG exchange, LONG comparand)
{
    return __sync_val_compare_and_swap(dst, comparand, exchange);
}
static inline LONG InterlockedIncrement(LONG volatile* p) { return __sync_add_and_fetch(p, 1); }
static inline void* _aligned_malloc(size_t size, size_t align)
{
    void* p = 0;
    if (posix_memalign(&p, align < sizeof(void*) ? sizeof(void*) : align, size)) return 0;
    return p;
}
static inline void _aligned_free(void* p) { free(p); }
#endif
