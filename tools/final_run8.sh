#!/bin/bash
# N-GPU lines of the round's final code: multi-GPU tests, bench on configs 2 / 4 / 5 (config 5 at its stated 4096 spp), the
# one-process multi-GPU handle.     usage: final_run8.sh N [tag]
N=$1; TAG=${2:-r02b_final}
O=gpurun_out/$TAG; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -2 > $O/multi_tests_g$N.log
$TR bench.py --gpus $N --steps 5 --warmup 3 > $O/bench_g$N.json 2> $O/bench_g$N.err
$TR bench.py --gpus $N --config 4 --steps 2 --warmup 1 > $O/bench_cfg4_g$N.json 2> $O/bench_cfg4_g$N.err
$TR bench.py --gpus $N --config 5 --steps 1 --warmup 1 --lean > $O/bench_cfg5_g$N.json 2> $O/bench_cfg5_g$N.err
python tools/multi_handle_run.py $N monkey_cfg2:1920:1080:256 cornell:1920:1080:256 > $O/multi_handle_g$N.jsonl 2> $O/multi_handle_g$N.err
tail -c 300 $O/*_g$N.json $O/*_g$N.jsonl $O/*_g$N.log 2>/dev/null
grep -l . $O/*_g$N.err 2>/dev/null | xargs -r -n1 tail -n 3
