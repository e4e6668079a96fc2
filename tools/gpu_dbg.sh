#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | grep "smoke\|worst" | tail -9
echo "--- gather upsample adjoint"; B200UNET_UPSAMPLE_BWD_GATHER=1 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | grep "smoke\|worst" | tail -5
echo "--- old small ops"; B200UNET_OLD_SMALL_OPS=1 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | grep "smoke\|worst" | tail -5
echo "--- kws 1"; B200UNET_HALO_KWS=1 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | grep "smoke\|worst" | tail -5
