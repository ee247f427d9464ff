#!/usr/bin/env bash
# First GPU call of a round: run the kernels that were written without a GPU, each under its own
# timeout (an unvalidated mbarrier pipeline can hang), then time the NeRF step with and without the
# tcgen05 dense layers. Everything lands in gpurun_out/.
#
#   gpurun --timeout 900 -- 'bash scripts/validate_unverified.sh'
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout 300 python -m pytest tests/test_zz_gpu_rays.py -x -q > gpurun_out/unverified_rays.log 2>&1
echo "rays: exit $?" | tee -a gpurun_out/unverified_summary.txt
ATMONR_RUN_UNVERIFIED=1 timeout 300 python -m pytest tests/test_zz_gpu_unverified_modes.py -q \
  > gpurun_out/unverified_modes.log 2>&1
echo "nerf modes: exit $?" | tee -a gpurun_out/unverified_summary.txt
ATMONR_RUN_UNVERIFIED=1 timeout 300 python -m pytest tests/test_zz_gpu_linear_tc.py -x -q \
  > gpurun_out/unverified_linear_tc.log 2>&1
rc=$?
echo "linear_tc: exit $rc" | tee -a gpurun_out/unverified_summary.txt
if [ $rc -eq 0 ]; then
  for tc in 0 1; do
    ATMONR_NERF_TC=$tc timeout 300 python scripts/bench_nerf.py > gpurun_out/nerf_tc$tc.json 2> gpurun_out/nerf_tc$tc.err
    echo "nerf tc=$tc: exit $? $(cat gpurun_out/nerf_tc$tc.json)" | tee -a gpurun_out/unverified_summary.txt
  done
fi
