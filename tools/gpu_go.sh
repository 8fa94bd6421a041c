#!/bin/bash
# build locally, then run tools/gpu_quick.sh on a B200 box — stops if the build fails
tag=${1:-quick}
python -c "import __graft_entry__ as g; g.build_cuda()" 2>&1 | grep -E "error|warning" | cut -c1-300 | head -20
if ! python -c "import __graft_entry__ as g; g.build_cuda()" >/dev/null 2>&1; then echo "BUILD FAILED"; exit 1; fi
/usr/local/graft/bin/gpurun --timeout 900 -- "bash tools/gpu_quick.sh $tag" 2>&1 | tail -15
