/* Hand-written config.h for the out-of-tree oracle build of the reference
 * FreeFEM 4.15 (/root/reference).  Test infrastructure only; not shipped in the
 * product.  The reference's own ./configure cannot run here (needs m4, bison,
 * flex, gfortran), so only the macros the core actually tests are set.
 * Deliberately NOT set: HAVE_LIBUMFPACK, HAVE_LIBARPACK, HAVE_HDF5,
 * HAVE_CBLAS_H, HAVE_MKL, anything MPI. */
#ifndef FFCUDA_ORACLE_CONFIG_H
#define FFCUDA_ORACLE_CONFIG_H
#define HAVE_DLFCN_H 1
#define HAVE_GETENV 1
#define HAVE_GETTIMEOFDAY 1
#define HAVE_STDINT_H 1
#define HAVE_STDLIB_H 1
#define HAVE_STRING_H 1
#define HAVE_SYS_TIME_H 1
#define HAVE_TIMES 1
#define HAVE_UNISTD_H 1
#define HAVE_CSTDDEF 1
#define HAVE_STDDEF_H 1
#define HAVE_ACOSH 1
#define HAVE_ASINH 1
#define HAVE_ATANH 1
#define HAVE_ERFC 1
#define HAVE_TGAMMA 1
#define HAVE_JN 1
#define HAVE_LIBDL 1
#define HAVE_LIBM 1
#define HAVE_REGEX_H 1
#define HAVE_SYSCONF 1
#define HAVE_SYS_MMAN_H 1
#define HAVE_SEMAPHORE_H 1
#define PACKAGE "FreeFEM"
#define PACKAGE_NAME "FreeFEM"
#define PACKAGE_VERSION "4.15"
#define VERSION "4.15"
#define PACKAGE_STRING "FreeFEM 4.15"
#define VersionFreeFem 4.15
#define VersionFreeFemDate "oracle hand build"
#define FF_PREFIX_DIR "/nonexistent-ff-prefix"
#endif
