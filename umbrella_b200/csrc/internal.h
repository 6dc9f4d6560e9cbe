// Library-internal host helpers shared by the translation units of libumbrella_b200.so.
#pragma once
#include "../../include/umbrella_b200.h"

namespace umb {
// every host entry point reports the kernels it enqueued (umb_launch_count)
void count_launches(unsigned n);
// number of SMs of the current device (cached)
int sm_count();
}  // namespace umb
