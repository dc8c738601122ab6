/* config.h for the oniguruma sources vendored in the reference tree (src/Utils/oniguruma), which its CMakeLists.txt generates at configure time
 * (src/Utils/oniguruma/CMakeLists.txt:22-41).  Hand-written equivalent for x86-64 Linux so that oracle/Makefile can compile those sources where
 * they lie without running cmake.  TEST INFRASTRUCTURE ONLY (oracle/_ref/libkoifish_reftok.so). */
#ifndef CONFIG_H
#define CONFIG_H
#define HAVE_ALLOCA_H 1
#define HAVE_STDINT_H 1
#define HAVE_SYS_TIMES_H 1
#define HAVE_SYS_TIME_H 1
#define HAVE_SYS_TYPES_H 1
#define HAVE_UNISTD_H 1
#define HAVE_INTTYPES_H 1
#define SIZEOF_INT 4
#define SIZEOF_LONG 8
#define SIZEOF_LONG_LONG 8
#define SIZEOF_VOIDP 8
#define PACKAGE "onig"
#define PACKAGE_VERSION "6.9.10"
#define VERSION "6.9.10"
#endif
