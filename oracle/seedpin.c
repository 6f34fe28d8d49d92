/*
 * oracle/seedpin.c -- TEST INFRASTRUCTURE ONLY.
 * The reference derives its sampling seed from time(NULL) and getpid()
 * (kbbq.cc:268-270; minion.hpp:377-397) and has no flag to override it.
 * Linking this object into the oracle CLI pins both, so two runs give the
 * same "Seed:" line and byte-identical output.
 * KBBQ_REF_TIME / KBBQ_REF_PID override the pinned values.
 */
#include <stdlib.h>
#include <time.h>
#include <unistd.h>
#include <sys/types.h>

time_t time(time_t *t) {
    const char *e = getenv("KBBQ_REF_TIME");
    time_t v = e ? (time_t)strtoll(e, NULL, 10) : (time_t)1700000000;
    if (t) *t = v;
    return v;
}

pid_t getpid(void) {
    const char *e = getenv("KBBQ_REF_PID");
    return e ? (pid_t)atoi(e) : (pid_t)4242;
}
