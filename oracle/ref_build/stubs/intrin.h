/* intrin.h stub -- TEST INFRASTRUCTURE (see windows.h stub). */
#ifndef ORACLE_STUB_INTRIN_H
#define ORACLE_STUB_INTRIN_H
#include <pthread.h>
#define _WriteBarrier() __asm__ __volatile__("" ::: "memory")
#define _ReadWriteBarrier() __asm__ __volatile__("" ::: "memory")
static inline unsigned long __threadid(void) { return (unsigned long)pthread_self(); }
#endif
