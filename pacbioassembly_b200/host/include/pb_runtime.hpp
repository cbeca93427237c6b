// pb_runtime.hpp -- process-wide GPU context for the header-level drop-in classes.
//
// The reference is single-threaded with global state (locator.cpp:26-30, spaced_seed.cpp:70-96); the drop-in
// classes mirror that with one lazily created pb_ctx per process.  Every call goes to the GPU through the C ABI
// of include/pacbio_b200.h; a failure (including "no device") aborts loudly -- there is no CPU path to fall to.
#pragma once

#include <cstdio>
#include <cstdlib>

#include "pacbio_b200.h"

namespace pb {

inline pb_ctx *&ctx_slot()
{
    static pb_ctx *c = nullptr;
    return c;
}

inline pb_ctx *ctx()
{
    pb_ctx *&c = ctx_slot();
    if (!c) {
        const char *dev = std::getenv("PB_DEVICE");
        int rc = pb_ctx_create(dev ? std::atoi(dev) : 0, &c);
        if (rc != PB_OK) {
            std::fprintf(stderr, "pacbio_b200: %s\n", pb_last_error(nullptr));
            std::abort();
        }
    }
    return c;
}

inline void check(int rc, const char *what)
{
    if (rc != PB_OK) {
        std::fprintf(stderr, "pacbio_b200: %s failed (%d): %s\n", what, rc, pb_last_error(ctx_slot()));
        std::abort();
    }
}

inline void shutdown()
{
    pb_ctx *&c = ctx_slot();
    if (c) pb_ctx_destroy(c);
    c = nullptr;
}

} // namespace pb
