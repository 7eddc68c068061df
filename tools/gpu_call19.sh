# (whatever GPU seconds are left) the rolled-plane-loop build vs the default build, cfg2 batches, same box
mkdir -p gpurun_out
timeout 40 python tools/roll_planes.py localrf_b200/csrc/liblrf_b200_roll.so roll 2>&1 | tail -2
timeout 40 python tools/roll_planes.py localrf_b200/csrc/liblrf_b200.so default 2>&1 | tail -2
