/* Hand-written stand-in for the file the reference's CMake generates from
 * include/kangaroo/config.h.in:1-33.  Only used to build oracle/_ref (the
 * UNMODIFIED reference kernels, compiled from /root/reference where they lie). */
#ifndef KANGAROO_CONFIG_H
#define KANGAROO_CONFIG_H
#define _UNIX_
#define _LINUX_
#define _GCC_
#define HAVE_THRUST
#define HAVE_NPP
#define CUDA_VERSION_MAJOR 12
#define CUDA_VERSION_MINOR 9
#if (__cplusplus > 199711L)
#define CALLEE_HAS_CPP11
#define CALLEE_HAS_RVALREF
#endif
#endif
