/*
 * config.h for compiling the UNMODIFIED reference sources (under /root/reference)
 * in place, without running the reference's autotools build.  Test infrastructure
 * only (see oracle/README.md).  Values match what the reference's own `configure`
 * detects on this x86-64 glibc image: in particular
 * HAVE_LOG2 => glibc log2(), not the log(x)/0.693 fallback of lib/misc.c:368-378.
 */
#ifndef ORACLE_REF_CONFIG_H
#define ORACLE_REF_CONFIG_H
#define HAVE_VPRINTF 1
#define STDC_HEADERS 1
#define X_DISPLAY_MISSING 1
#define HAVE_SIGNED_SHIFT 1
#define SIZEOF_CHAR 1
#define SIZEOF_INT 4
#define SIZEOF_SHORT 2
#define HAVE_LOG2 1
#define HAVE_MEMMOVE 1
#define HAVE_STRCASECMP 1
#define HAVE_STRDUP 1
#define HAVE_ASSERT_H 1
#define HAVE_FEATURES_H 1
#define HAVE_SETJMP_H 1
#define HAVE_STRING_H 1
#define HAVE_UNISTD_H 1
#define HAVE_LIBM 1
#define PACKAGE "fiasco"
#define VERSION "1.3"
#ifndef FIASCO_SHARE
#define FIASCO_SHARE "/nonexistent/share/fiasco/"
#endif
#endif
