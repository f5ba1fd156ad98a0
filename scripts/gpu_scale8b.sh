# 8-GPU check of the final round-2 build (one box): weak-scaling bench line and the C++ multi-GPU host
O=gpurun_out/${OUT:-r2b_n8}; mkdir -p $O
N=${N:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
nvidia-smi -L | wc -l > $O/ngpu.txt
timeout 400 $TR bench.py --gpus $N --steps 10 --warmup 3 > $O/bench_n$N.json 2> $O/bench_n$N.err
tail -c 700 $O/bench_n$N.json
python scripts/make_event_files.py sipm8x8_scint $((12500000 * N)) /tmp/ev8 > $O/cxx_files.txt 2>&1
timeout 300 eic-opticks_b200/apps/PhoxMultiGPU -g /tmp/ev8/geom -G /tmp/ev8/gs.npy --gpus $N --events 10 --max-bounce 32 > $O/cxx_multigpu_n$N.txt 2>&1
cat $O/cxx_multigpu_n$N.txt
