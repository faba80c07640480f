#!/bin/bash
# Turn the artefacts of tools/r2_profile.sh (gpurun_out/r2z_*) into the committed evidence under profiles/ (tag r2).
set -e
cd "$(dirname "$0")/.."
DIGEST=$(cat monte-carlo-portfolio_b200/build/stamp)
python tools/ncu_summary.py gpurun_out/r2z_prof.ncu-rep r2 gpurun_out/r2z_bench_launches.csv
python tools/ncu_figures.py gpurun_out/r2z_prof.ncu-rep "$DIGEST" r2
python tools/ncu_hot.py gpurun_out/r2z_prof.ncu-rep 25 'small_sweep_packed' > profiles/r2_sweep_hotspots.txt
python tools/ncu_hot.py gpurun_out/r2z_prof.ncu-rep 25 path_kernel_tc 1 > profiles/r2_paths_tc_hotspots.txt
python tools/ncu_hot.py gpurun_out/r2z_prof.ncu-rep 25 large_sweep_tc 0 > profiles/r2_tc_hotspots.txt
python tools/ncu_hot.py gpurun_out/r2z_prof.ncu-rep 25 large_sweep_tc 6 > profiles/r2_tc_bounded_hotspots.txt
python tools/sass_counts.py r2
cp gpurun_out/r2z_plain.log profiles/r2_profile_target_plain.log
echo "collected for digest $DIGEST"
