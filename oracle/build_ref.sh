#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY.
# Compiles the reference's own CPU sources WHERE THEY LIE under $GST_REFERENCE
# (default /root/reference, read-only) together with oracle/ref_glue.cpp into
# oracle/_ref/libgst_ref.so.  No reference source is copied into this repo; the
# only outputs are objects + the .so under oracle/_ref/ (git-ignored, shipped to
# the GPU box by gpurun like any other built artefact).
#
# The reference's CMake build is NOT used (it needs OpenCL + GL); this is the
# direct g++ recipe SURVEY.md section 8(c) found to work: encoder + CPU rANS + CPU
# wavelet compile unmodified with -std=c++11 -fms-extensions once a tiny shim
# gpu.h (oracle/ref_shim/) satisfies ans/ans_ocl.h's include.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${GST_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -d "$REF/codec" ]; then
  echo "build_ref.sh: $REF not present; keeping any prebuilt $OUT/libgst_ref.so" >&2
  exit 0
fi
mkdir -p "$OUT/obj"
CXXFLAGS="-std=c++11 -fms-extensions -O2 -DNDEBUG -fPIC -w"
INC="-I$HERE/ref_shim -I$REF/codec -I$REF/ans -I$REF/lib/include -I$REF/lib/vptree/include -I$REF/lib -I$REF/lib/vptree/src"
CPP_SRCS="ans/encode.cpp ans/decode.cpp ans/histogram.cpp ans/ans_ocl_encode.cpp
  codec/wavelet.cpp codec/data_stream.cpp codec/codec_base.cpp codec/image.cpp
  codec/image_processing.cpp codec/image_utils.cpp codec/dxt_image.cpp codec/encoder.cpp
  codec/entropy.cpp lib/vptree/src/vptree_cpp.cc"
C_SRCS="lib/vptree/src/vptree.c lib/vptree/src/geom.c lib/vptree/src/pqueue.c"
OBJS=""
pids=""
for s in $CPP_SRCS; do
  o="$OUT/obj/$(echo "$s" | tr '/' '_').o"
  OBJS="$OBJS $o"
  if [ ! -f "$o" ] || [ "$REF/$s" -nt "$o" ]; then
    g++ $CXXFLAGS $INC -c "$REF/$s" -o "$o" &
    pids="$pids $!"
  fi
done
for s in $C_SRCS; do
  o="$OUT/obj/$(echo "$s" | tr '/' '_').o"
  OBJS="$OBJS $o"
  if [ ! -f "$o" ] || [ "$REF/$s" -nt "$o" ]; then
    gcc -O2 -DNDEBUG -fPIC -w -I"$REF/lib/vptree/src" -I"$REF/lib/vptree/include" -c "$REF/$s" -o "$o" &
    pids="$pids $!"
  fi
done
for p in $pids; do wait "$p"; done
g++ $CXXFLAGS $INC -c "$HERE/ref_glue.cpp" -o "$OUT/obj/ref_glue.o"
g++ -shared -o "$OUT/libgst_ref.so" "$OUT/obj/ref_glue.o" $OBJS -lpthread
echo "built $OUT/libgst_ref.so"
