/* Minimal config.h for compiling the UNMODIFIED liba52 sources where they lie
 * under /root/reference (oracle/Makefile).  Test infrastructure only.
 * Float samples (no LIBA52_FIXED / LIBA52_DOUBLE), no djbfft, plain malloc -
 * the same configuration the reference's shipped configure picks on x86-64
 * and the one vc++/config.h uses for the ACM build. */
#ifndef ORACLE_REFBUILD_CONFIG_H
#define ORACLE_REFBUILD_CONFIG_H
#define HAVE_INTTYPES_H 1
#define HAVE_STDINT_H 1
#define HAVE_STDLIB_H 1
#define HAVE_STRING_H 1
#define PACKAGE "a52dec"
#define VERSION "0.7.5-cvs"
#endif
