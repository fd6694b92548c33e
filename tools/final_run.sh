#!/bin/bash
# Final numbers of a round on N GPUs: bench (config 2, both arms at N = 1), the other configs, the multi-GPU handle, smoke().
# usage: final_run.sh N [tag]      results under gpurun_out/<tag>/
N=$1; TAG=${2:-r02_final}
O=gpurun_out/$TAG; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
if [ "$N" = "1" ]; then TR="python"; fi
$TR bench.py --gpus $N --steps 5 --warmup 3 > $O/bench_g$N.json 2> $O/bench_g$N.err
if [ "$N" = "1" ]; then
  python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref_g1.json 2> $O/bench_ref_g1.err
  python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1
  python bench.py --config cornell1080 --steps 5 --warmup 3 --no-extra > $O/bench_cornell_g1.json 2> $O/bench_cornell_g1.err
  python bench.py --config 1 --steps 5 --warmup 3 --no-extra > $O/bench_cfg1_g1.json 2> $O/bench_cfg1_g1.err
  python bench.py --config 3 --steps 3 --warmup 2 --no-extra > $O/bench_cfg3_g1.json 2> $O/bench_cfg3_g1.err
  python bench.py --config 4 --steps 2 --warmup 1 --no-extra > $O/bench_cfg4_g1.json 2> $O/bench_cfg4_g1.err
else
  $TR bench.py --gpus $N --config 4 --steps 2 --warmup 1 > $O/bench_cfg4_g$N.json 2> $O/bench_cfg4_g$N.err
  python tools/multi_handle_run.py $N monkey_cfg2:1920:1080:256 cornell:1920:1080:256 > $O/multi_handle_g$N.jsonl 2> $O/multi_handle_g$N.err
fi
tail -c 400 $O/*_g$N.json* $O/smoke.log 2>/dev/null
grep -l . $O/*_g$N.err 2>/dev/null | xargs -r -n1 tail -n 3
