# full-set captures of the kernels added late in round 2 (final versions): single-layer walk, two-layer walk, dynamic
# two-layer walk, padded-voxel kernel.  Each command runs once without ncu first.
set -x
python profiles/tune_stage.py auto 16 > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_pillar_walk -s 10 -c 1 -f -o gpurun_out/r02f_walk python profiles/tune_stage.py auto 16 > gpurun_out/ncu.log 2>&1
python profiles/scripts/stack_times.py > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_pillar_walk -s 8 -c 1 -f -o gpurun_out/r02f_walk_two python profiles/scripts/stack_times.py > gpurun_out/ncu.log 2>&1
python profiles/scripts/dyn_times.py > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_pillar_walk -s 20 -c 1 -f -o gpurun_out/r02f_walk_dyn2 python profiles/scripts/dyn_times.py > gpurun_out/ncu.log 2>&1
python profiles/scripts/padded_times.py > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_pfn_padded -s 3 -c 1 -f -o gpurun_out/r02f_padded python profiles/scripts/padded_times.py > gpurun_out/ncu.log 2>&1
ls -la gpurun_out/r02f_*.ncu-rep
