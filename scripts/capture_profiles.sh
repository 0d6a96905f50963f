#!/bin/bash
# Run on the GPU box (one GPU): bench line, ncu launch list and --set full captures of the hot kernels.
#   gpurun --timeout 1200 -- 'bash scripts/capture_profiles.sh r01'
# Outputs land in gpurun_out/; scripts/refresh_profiles.py turns them into profiles/<tag>_*.
TAG=${1:-r02}
O=gpurun_out
mkdir -p $O
python bench.py --steps 10 --warmup 3 > $O/${TAG}_bench_n1.json 2> $O/${TAG}_bench.err
tail -c 2500 $O/${TAG}_bench_n1.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${TAG}_launches.csv \
    python scripts/profile_frame.py 2 > $O/${TAG}_launches.log 2>&1
# QUICK=1 skips the two big captures (reuse them when only the small kernels changed)
for k in $([ -n "$QUICK" ] || echo k_trace k_shade); do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 10 -c 2 -o $O/${TAG}_prof_$k -f \
      python scripts/profile_frame.py 1 > $O/${TAG}_prof_$k.log 2>&1
done
# DRAM bytes of EVERY k_trace / k_shade launch of one whole frame (roofline.traffic = mean per launch over the frame, comparable with
# roofline.achieved which is the mean over the same launches)
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k 'regex:k_trace|k_shade' -c 400 --csv \
    --log-file $O/${TAG}_frame_traffic.csv python scripts/profile_frame.py 1 > $O/${TAG}_frame_traffic.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:k_reproject|k_denoise|k_primary|k_tail$' -c 4 -o $O/${TAG}_prof_misc -f \
    python scripts/profile_frame.py 1 > $O/${TAG}_prof_misc.log 2>&1
ls -la $O | tail -12
