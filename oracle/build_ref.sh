#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY -- builds the UNMODIFIED reference (stivalaa/cuda_satabsearch)
# from the sources where they lie under /root/reference into oracle/_ref/ (git-ignored,
# but shipped to the GPU box by gpurun).  Nothing from the reference is copied into the
# repository; only compiled objects/binaries land in oracle/_ref/.
#
# Follows the reference's own Makefile rules (nvcc_src_current/Makefile:93-116):
#   * cudaSaTabsearch_host.o   : g++ -x c++ of the kernel source (the `-c` CPU path)
#   * three nvcc builds of the kernel source (-DCUDA [-DUSE_SHARED_MEMORY|-DSMALL_MAXDIM])
#   * nvcc build of the host driver, g++ builds of the parser and the Gumbel statistics
# plus `-gencode arch=compute_100a,code=sm_100a` (the Makefile passes no -arch at all)
# and the three stub headers in oracle/ref_stubs/ that stand in for the CUDA-samples
# headers this image does not have.
#
# Outputs:
#   oracle/_ref/cudaSaTabsearch_ref        reference binary, MAXDIM_GPU 96 (as shipped)
#   oracle/_ref/cudaSaTabsearch_ref_md32   same sources with saparams.h:18 MAXDIM_GPU 32 (the 2013
#                                          value) -- reproduces old/nvcc_src_cuda5/cpu_cudaSaTabsearch.o1462445
#                                          The one-line edit is made on a scratch copy under $TMPDIR.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${SATS_REFERENCE_DIR:-/root/reference}"
SRC="$REF/nvcc_src_current"
OUT="$HERE/_ref"
STUBS="$HERE/ref_stubs"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
CUDA_INC="${CUDA_INC:-/usr/local/cuda/include}"
ARCH="-gencode arch=compute_100a,code=sm_100a"

if [ ! -d "$SRC" ]; then
  echo "build_ref.sh: $SRC not present (GPU box?) -- keeping prebuilt oracle/_ref" >&2
  exit 0
fi
mkdir -p "$OUT"

build_variant() {  # $1 = source dir, $2 = output binary, $3 = object dir
  local S="$1" BIN="$2" O="$3"
  mkdir -p "$O"
  g++ -x c++ -c -O3 -I"$CUDA_INC" -I"$S" -o "$O/host_kernel.o" "$S/cudaSaTabsearch_kernel.cu"
  "$NVCC" -c -O3 --use_fast_math $ARCH -Wno-deprecated-gpu-targets -I"$STUBS" -I"$S" \
      -DCUDA -DUSE_SHARED_MEMORY -o "$O/kernel.o" "$S/cudaSaTabsearch_kernel.cu" 2>/dev/null
  "$NVCC" -c -O3 --use_fast_math $ARCH -Wno-deprecated-gpu-targets -I"$STUBS" -I"$S" \
      -DCUDA -o "$O/kernel_noshared.o" "$S/cudaSaTabsearch_kernel.cu" 2>/dev/null
  "$NVCC" -c -O3 --use_fast_math $ARCH -Wno-deprecated-gpu-targets -I"$STUBS" -I"$S" \
      -DCUDA -DSMALL_MAXDIM -o "$O/kernel_noshared_small.o" "$S/cudaSaTabsearch_kernel.cu" 2>/dev/null
  "$NVCC" -c -O3 --use_fast_math $ARCH -Wno-deprecated-gpu-targets -I"$STUBS" -I"$S" \
      -DCUDA -o "$O/main.o" "$S/cudaSaTabsearch.cu" 2>/dev/null
  g++ -c -O3 -I"$S" -o "$O/parsetableaux.o" "$S/parsetableaux.c" 2>/dev/null
  g++ -c -O3 -I"$S" -o "$O/gumbelstats.o" "$S/gumbelstats.c"
  "$NVCC" $ARCH -o "$BIN" "$O"/host_kernel.o "$O"/kernel.o "$O"/kernel_noshared.o \
      "$O"/kernel_noshared_small.o "$O"/main.o "$O"/parsetableaux.o "$O"/gumbelstats.o -lm
}

build_variant "$SRC" "$OUT/cudaSaTabsearch_ref" "$OUT/obj96"

# MAXDIM_GPU 32 variant (2013 golden).  Scratch copy outside the repo; only the binary is kept.
SCR="$(mktemp -d "${TMPDIR:-/tmp}/sats_ref32.XXXXXX")"
trap 'rm -rf "$SCR"' EXIT
cp "$SRC"/*.cu "$SRC"/*.c "$SRC"/*.h "$SCR"/
sed -i 's/^#define MAXDIM_GPU 96 /#define MAXDIM_GPU 32 /' "$SCR/saparams.h"
grep -q '^#define MAXDIM_GPU 32 ' "$SCR/saparams.h"
build_variant "$SCR" "$OUT/cudaSaTabsearch_ref_md32" "$SCR/obj32"
echo "build_ref.sh: built $OUT/cudaSaTabsearch_ref and $OUT/cudaSaTabsearch_ref_md32"
