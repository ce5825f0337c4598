// rt_internal.h -- shared between the host-side translation units of librt_b200.so.
#ifndef RT_INTERNAL_H
#define RT_INTERNAL_H

#include <string>

#include "rt_b200.h"

// Records the message for rt_last_error() (thread-local) and returns `code`.
int rt_fail(int code, const std::string &msg);

#endif
