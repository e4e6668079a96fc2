#!/bin/bash
# A/B copies of the library for compile-time kernel variants (loaded with B200UNET_LIB=<path>).  None at the moment: the
# round-2 variants became runtime switches (B200UNET_HALO_KWS, B200UNET_HALO_1X1_DENSE, B200UNET_S2_ZERO_INSERT,
# B200UNET_OLD_SMALL_OPS, B200UNET_TILED_UPSAMPLE_BWD); the Makefile keeps BUILD= / OUT= / EXTRA= for the next one.
cd "$(dirname "$0")/../3dunetcnn_b200/csrc"
make -j8 2>&1 | grep -E "error|warning"
ls -la ../*.so
