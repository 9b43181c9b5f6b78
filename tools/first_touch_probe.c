/* first_touch_probe.c — how fast can this host populate a fresh anonymous allocation?  (What a fresh
 * Vector{Float64}(undef, n) destination costs the *_host calls.)  gcc -O2 -pthread first_touch_probe.c
 * usage: probe <GiB> <threads> <mode>   mode 0: touch one byte per 4 KiB; 1: MADV_POPULATE_WRITE per slice;
 * 2: memset the slice.  The region is madvise(MADV_HUGEPAGE)'d when the 4th argument is 1. */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <time.h>
#ifndef MADV_POPULATE_WRITE
#define MADV_POPULATE_WRITE 23
#endif
static char* base; static size_t bytes; static int nthreads, mode;
static double now(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }
static void* work(void* arg) {
    long id = (long)arg;
    size_t per = (bytes / nthreads) & ~((size_t)(2u << 20) - 1);
    char* p = base + id * per; size_t len = id == nthreads - 1 ? bytes - id * per : per;
    if (mode == 1) { if (madvise(p, len, MADV_POPULATE_WRITE) != 0) perror("populate"); }
    else if (mode == 2) memset(p, 1, len);
    else for (size_t i = 0; i < len; i += 4096) p[i] = 1;
    return 0;
}
int main(int argc, char** argv) {
    double gib = argc > 1 ? atof(argv[1]) : 4; nthreads = argc > 2 ? atoi(argv[2]) : 16; mode = argc > 3 ? atoi(argv[3]) : 0;
    int huge = argc > 4 ? atoi(argv[4]) : 1;
    bytes = (size_t)(gib * (1ull << 30));
    base = mmap(0, bytes + (2u << 20), PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (base == MAP_FAILED) { perror("mmap"); return 1; }
    base = (char*)(((size_t)base + (2u << 20) - 1) & ~((size_t)(2u << 20) - 1));
    if (huge) madvise(base, bytes, MADV_HUGEPAGE);
    pthread_t th[256]; double t0 = now();
    for (long i = 0; i < nthreads; ++i) pthread_create(&th[i], 0, work, (void*)i);
    for (int i = 0; i < nthreads; ++i) pthread_join(th[i], 0);
    double dt = now() - t0;
    char line[256]; long ahp = -1; FILE* f = fopen("/proc/self/smaps_rollup", "r");
    if (f) { while (fgets(line, sizeof line, f)) if (!strncmp(line, "AnonHugePages:", 14)) ahp = atol(line + 14); fclose(f); }
    printf("%.0f GiB threads %3d mode %d huge %d: %.3f s = %6.1f GB/s  AnonHugePages %ld MiB\n", gib, nthreads, mode, huge, dt, bytes / dt / 1e9, ahp / 1024);
    return 0;
}
