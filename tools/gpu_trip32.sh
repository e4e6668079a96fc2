#!/bin/bash
cd "$(dirname "$0")/.."
for f in 0 1 2; do echo "--- FLAGS=$f"; B200UNET_HALO_FLAGS=$f timeout 600 python tools/conv_determinism.py 4 2>&1 | grep "\[det\] conv" | grep -E "ci(32|64) +co(32|128) +r(128|32) +(plain|mode1)"; done
