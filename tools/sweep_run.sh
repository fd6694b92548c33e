#!/bin/bash
# One allocation of N GPUs: multi-GPU tests, the bench on configs 2 / 4 / 5 at their stated spp, the multi-GPU handle.
# usage: sweep_run.sh N [cfg5_spp]      results under gpurun_out/r02_sweep/
N=$1; SPP5=${2:-4096}
O=gpurun_out/r02_sweep; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
if [ "$N" = "1" ]; then TR="python"; fi
python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -2 > $O/multi_tests_g$N.log
$TR bench.py --gpus $N --steps 3 --warmup 3 > $O/bench_cfg2_g$N.json 2> $O/bench_cfg2_g$N.err
$TR bench.py --gpus $N --config 4 --steps 2 --warmup 1 > $O/bench_cfg4_g$N.json 2> $O/bench_cfg4_g$N.err
$TR bench.py --gpus $N --config 5 --spp $SPP5 --steps 1 --warmup 1 --lean > $O/bench_cfg5_g$N.json 2> $O/bench_cfg5_g$N.err
if [ "$N" != "1" ]; then
  python tools/multi_handle_run.py $N monkey_cfg2:1920:1080:256 cornell:1920:1080:256 serre:3840:2160:64 > $O/multi_handle_g$N.jsonl 2> $O/multi_handle_g$N.err
fi
tail -c 600 $O/*_g$N.json $O/*_g$N.jsonl $O/*_g$N.log 2>/dev/null
grep -l . $O/*_g$N.err 2>/dev/null | xargs -r -n1 tail -n 3
