// vb_stubs.cu -- entry points declared in include/vecb200.h whose kernels are not built yet.
// They fail loudly (VB_ESTATE); there is no CPU fallback behind any of them.
#include "vb_common.cuh"

#define NOT_YET(name)                                          \
    do {                                                       \
        vb::set_error(name ": not implemented in this build"); \
        return VB_ESTATE;                                      \
    } while (0)

extern "C" {
}
