#!/bin/sh
# oracle/_ref: the one piece of the reference's hot path that compiles here from its own source — the body of the legacy 802.11b transmit
# filter, FIR37SSE_INTRINSIC (kernel/bb/dot11b/bbb_fir.c:390-566, pure SSE2 intrinsics) with its coefficient table (bbb_fir.c:21-63).
# The rest of that file needs the Windows kernel headers (KSPIN_LOCK, KIRQL, MSVC inline assembly), so the two spans are cut out of the
# source WHERE IT LIES, joined with a few typedefs and a plain-C entry point, and compiled into oracle/_ref/ (git-ignored; it travels to the
# GPU box with the snapshot).  Nothing of the reference is copied into the repository.  TEST INFRASTRUCTURE ONLY.
#   usage: oracle/build_ref.sh [reference root, default /root/reference]
set -e
REF="${1:-/root/reference}"
SRC="$REF/kernel/bb/dot11b/bbb_fir.c"
HERE="$(cd "$(dirname "$0")" && pwd)"
[ -f "$SRC" ] || { echo "build_ref: $SRC not found (fine outside the build container: the prebuilt oracle/_ref is used)"; exit 0; }
mkdir -p "$HERE/_ref"
OUT="$HERE/_ref/fir37_ref.c"
{
  cat <<'HDR'
#include <emmintrin.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
typedef long HRESULT; typedef short SHORT; typedef short* PSHORT; typedef unsigned int DWORD; typedef uintptr_t UPOINTER;
typedef struct { signed char re, im; } COMPLEX8, *PCOMPLEX8;
#define S_OK 0
#define E_FAIL (-1)
#define FALSE 0
#define A16 __attribute__((aligned(16)))
#define TX_FIR_DEPTH 37
HDR
  sed -n '/^const A16 SHORT SSEFilterTaps/,/^};/p' "$SRC"
  sed -n '/^#define __MULLO_ADDS/,$p' "$SRC"
  cat <<'TAIL'
/* plain-C entry point: BB11BPMDSpreadFIR4SSE (bbb_fir.c:92-110) without the spin lock; src may be unaligned, samples past n_in read as zero */
void ref_fir37(const signed char* src, unsigned n_in, signed char* dst) {
    static A16 SHORT temp[(TX_FIR_DEPTH + 3) * 8];
    size_t bytes = (size_t)n_in * 2;
    signed char* a = (signed char*)aligned_alloc(16, bytes + 64), *o = (signed char*)aligned_alloc(16, bytes + 64);
    memset(a, 0, bytes + 64); memcpy(a, src, bytes);
    memset(temp, 0, sizeof temp);
    if (n_in >> 3) FIR37SSE_INTRINSIC((PCOMPLEX8)a, SSEFilterTaps[0], n_in >> 3, temp, (PCOMPLEX8)o);
    _mm_sfence();
    memcpy(dst, o, (size_t)(n_in >> 3) * 16);
    free(a); free(o);
}
TAIL
} > "$OUT"
${CC:-gcc} -O2 -msse2 -fPIC -shared -w -o "$HERE/_ref/libfir37_ref.so" "$OUT"
echo "build_ref: oracle/_ref/libfir37_ref.so"
