/* Force-included (-include) MSVC compatibility shim -- test infrastructure only.
 *
 * The reference is a Visual Studio project; Main.c:43,57, Network.c:29,129,171,
 * comparator.c:30-31 and ViT_seq.c:516 use MSVC-only names. Mapping them here
 * lets the reference sources compile unmodified with gcc.
 */
#ifndef VITB200_MSVC_COMPAT_H
#define VITB200_MSVC_COMPAT_H
#include <stdio.h>
#include <string.h>
#include <errno.h>
#include <time.h>
typedef int errno_t;
static inline errno_t vitb200_fopen_s(FILE **f, const char *name, const char *mode)
{
    *f = fopen(name, mode);
    return *f ? 0 : errno;
}
#define fopen_s vitb200_fopen_s
static inline errno_t vitb200_strncpy_s(char *dst, size_t dstsz, const char *src, size_t count)
{
    size_t n = count < dstsz - 1 ? count : dstsz - 1;
    memcpy(dst, src, n);
    dst[n] = '\0';
    return 0;
}
#define strncpy_s vitb200_strncpy_s
#ifndef CLK_TCK
#define CLK_TCK CLOCKS_PER_SEC
#endif
#endif
