/*
 * oracle/parfor.h -- TEST INFRASTRUCTURE ONLY.  Minimal pthread parallel-for (this image's gcc has
 * no libgomp).  Splits [0, total) into contiguous chunks over ORACLE_THREADS (default: all online
 * cores) threads.  Results are independent of the thread count: every item writes its own slot.
 */
#ifndef GEOM_ORACLE_PARFOR_H
#define GEOM_ORACLE_PARFOR_H
#include <pthread.h>
#include <stdlib.h>
#include <unistd.h>

typedef void (*parfor_fn)(long begin, long end, void *ctx);
typedef struct { parfor_fn fn; long begin, end; void *ctx; } parfor_task;

static void *parfor_tramp(void *p)
{
    parfor_task *t = (parfor_task *)p;
    t->fn(t->begin, t->end, t->ctx);
    return 0;
}

static int parfor_threads(void)
{
    const char *e = getenv("ORACLE_THREADS");
    long t = e ? atol(e) : sysconf(_SC_NPROCESSORS_ONLN);
    if (t < 1) t = 1;
    if (t > 256) t = 256;
    return (int)t;
}

static void parfor(long total, parfor_fn fn, void *ctx)
{
    int nt = parfor_threads();
    if (total < 64 || nt == 1) { fn(0, total, ctx); return; }
    pthread_t th[256];
    parfor_task tk[256];
    long chunk = (total + nt - 1) / nt;
    int used = 0;
    for (int i = 0; i < nt; i++) {
        long b = i * chunk, e = b + chunk < total ? b + chunk : total;
        if (b >= e) break;
        tk[used].fn = fn; tk[used].begin = b; tk[used].end = e; tk[used].ctx = ctx;
        pthread_create(&th[used], 0, parfor_tramp, &tk[used]);
        used++;
    }
    for (int i = 0; i < used; i++) pthread_join(th[i], 0);
}
#endif
