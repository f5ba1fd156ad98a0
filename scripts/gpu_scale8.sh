# 8-GPU session of round 2 (one box): weak-scaling bench line, BASELINE config 4 at its stated size, the C++ multi-GPU host
O=gpurun_out/${OUT:-r2_n8}; mkdir -p $O
N=${N:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
nvidia-smi -L | wc -l > $O/ngpu.txt
timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 > $O/bench_n$N.json 2> $O/bench_n$N.err
tail -c 600 $O/bench_n$N.json
timeout 900 $TR bench.py --gpus $N --steps 3 --warmup 3 --workload pmt_wall_torch --photons 125000000 > $O/bench_pmt_wall_1G_n$N.json 2> $O/bench_pmt_wall_1G_n$N.err
tail -c 400 $O/bench_pmt_wall_1G_n$N.json; tail -3 $O/bench_pmt_wall_1G_n$N.err
python scripts/make_event_files.py sipm8x8_scint $((12500000 * N)) /tmp/ev8 > $O/cxx_files.txt 2>&1
timeout 600 eic-opticks_b200/apps/PhoxMultiGPU -g /tmp/ev8/geom -G /tmp/ev8/gs.npy --gpus $N --events 10 --max-bounce 32 > $O/cxx_multigpu_n$N.txt 2>&1
cat $O/cxx_multigpu_n$N.txt
timeout 300 eic-opticks_b200/apps/PhoxMultiGPU -g /tmp/ev8/geom -G /tmp/ev8/gs.npy --gpus 1 --events 2 --max-bounce 32 > $O/cxx_multigpu_n1_same_event.txt 2>&1
cat $O/cxx_multigpu_n1_same_event.txt
ls -la $O
