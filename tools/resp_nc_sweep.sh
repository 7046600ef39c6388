#!/bin/bash
# Build the library with the response kernel's columns-per-lane set to 3, 4 and 6 (5 is the default build) into
# golf_b200/_lib/libgolf_b200_nc<N>.so, for A/B timing on the GPU box:
#   GOLF_B200_SO=golf_b200/_lib/libgolf_b200_nc3.so python tools/step_events.py
set -e
cd "$(dirname "$0")/.."
for NC in "$@"; do
  OBJ=/tmp/golf_nc$NC; mkdir -p $OBJ
  for f in golf_b200/csrc/*.cu; do
    n=$(basename $f .cu)
    if [ $n = lpc_ss_mp ]; then
      for mp in 4 8 12 16 20 24 32 40; do
        nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-fvisibility=hidden --expt-relaxed-constexpr \
          -DGOLF_MP=$mp -DGOLF_RESP_NC=${NC%%w*} -DGOLF_RESP_WPB=${WPB:-1} ${RESDEF} -I include -c $f -o $OBJ/${n}$mp.o &
      done
    else
      nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-fvisibility=hidden --expt-relaxed-constexpr \
        -DGOLF_RESP_NC=${NC%%w*} -DGOLF_RESP_WPB=${WPB:-1} ${RESDEF} -I include -c $f -o $OBJ/$n.o &
    fi
  done
  wait
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o golf_b200/_lib/libgolf_b200_nc${NC}w${WPB:-1}${TAG}.so $OBJ/*.o -Xlinker --no-undefined -lcudart_static -ldl -lrt -lpthread
  echo built golf_b200/_lib/libgolf_b200_nc${NC}w${WPB:-1}${TAG}.so
done
