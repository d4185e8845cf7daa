#!/bin/bash
# scripts/mgpu_cli_check.sh N -- host/raytracer on 1 GPU and on N GPUs with each frame assembly of libskr_mgpu.so:
# row bands (default: pixels stored into the band owner's memory, N band copies), direct (SKR_MGPU_NO_BANDS=1: every GPU's
# kernel stores into the page-locked host frame), peer stores into GPU 0's frame
# (SKR_MGPU_NO_DIRECT=1) and NCCL all-gather (SKR_MGPU_NO_DIRECT=1 SKR_MGPU_NO_P2P=1).  The five PPM files must be
# byte-identical, deterministic, jittered and --gillum frames alike.
set -e
N=${1:-2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/mg
ok=1
for mode in "--width 1920 --height 1080 --jsample 3 --shadow --seed 5" "--width 640 --height 360 --gillum 8 --jsample 2 --shadow --seed 6" "--width 641 --height 357"; do
  host/raytracer --path tests/golden/tiny.scn --output gpurun_out/mg/a.ppm $mode > /dev/null
  host/raytracer --path tests/golden/tiny.scn --output gpurun_out/mg/b.ppm --gpus $N --stats $mode | tail -1 | cut -c1-200
  SKR_MGPU_NO_BANDS=1 host/raytracer --path tests/golden/tiny.scn --output gpurun_out/mg/e.ppm --gpus $N $mode > /dev/null
  SKR_MGPU_NO_DIRECT=1 host/raytracer --path tests/golden/tiny.scn --output gpurun_out/mg/c.ppm --gpus $N $mode > /dev/null
  SKR_MGPU_NO_DIRECT=1 SKR_MGPU_NO_P2P=1 host/raytracer --path tests/golden/tiny.scn --output gpurun_out/mg/d.ppm --gpus $N $mode > /dev/null
  if cmp -s gpurun_out/mg/a.ppm gpurun_out/mg/b.ppm && cmp -s gpurun_out/mg/a.ppm gpurun_out/mg/c.ppm && cmp -s gpurun_out/mg/a.ppm gpurun_out/mg/d.ppm && cmp -s gpurun_out/mg/a.ppm gpurun_out/mg/e.ppm; then echo "identical: $mode"; else echo "DIFFERENT: $mode"; ok=0; fi
done
[ $ok = 1 ] && echo "mgpu cli check OK ($N GPUs)"
