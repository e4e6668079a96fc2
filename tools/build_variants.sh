#!/bin/bash
# A/B copies of the library for compile-time kernel variants (loaded with B200UNET_LIB=<path>; see tools/gpu_experiments.sh)
cd "$(dirname "$0")/../3dunetcnn_b200/csrc"
make -j8 2>&1 | grep -E "error|warning"
make -j8 BUILD=build_kws1 OUT=../libb200unet_kws1.so EXTRA=-DB200_HALO_KWS1 2>&1 | grep -E "error|warning"   # one (kh,kw) box per weight stage (round-1 issue loop)
ls -la ../*.so
