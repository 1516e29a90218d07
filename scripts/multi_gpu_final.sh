#!/bin/bash
# Final multi-GPU evidence of the round on ONE 8-GPU box (gpurun --gpus 8 -- bash scripts/multi_gpu_final.sh [tag]):
# bench.py at N=8 and N=4 with the shipped kernels, and the serial-sum check of the gathered label stats (800-clip corpus at N=8 vs N=1).
TAG=${1:-r02n}
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 8 --master-port 29700 bench.py --gpus 8 --steps 10 --warmup 3 > $OUT/${TAG}_bench_n8.log 2> $OUT/${TAG}_bench_n8.err
echo "== bench n8 rc=$?"
timeout 300 $TR --nproc-per-node 4 --master-port 29704 bench.py --gpus 4 --steps 10 --warmup 3 > $OUT/${TAG}_bench_n4.log 2> $OUT/${TAG}_bench_n4.err
echo "== bench n4 rc=$?"
timeout 300 $TR --nproc-per-node 8 --master-port 29703 scripts/corpus_stream.py --clips 800 --mode vit --batch 50 --consume > $OUT/${TAG}_corpus_vit_800_n8.log 2> $OUT/${TAG}_corpus_vit_800_n8.err
timeout 300 python scripts/corpus_stream.py --clips 800 --mode vit --batch 50 --consume > $OUT/${TAG}_corpus_vit_800_n1.log 2> $OUT/${TAG}_corpus_vit_800_n1.err
echo "== serial-sum check"; tail -1 $OUT/${TAG}_corpus_vit_800_n8.log | cut -c1-330; tail -1 $OUT/${TAG}_corpus_vit_800_n1.log | cut -c1-330
python - <<PY
import json
for n in (8, 4):
    try:
        l = json.loads(open("$OUT/${TAG}_bench_n%d.log" % n).read().strip().splitlines()[-1])
        print(n, "value", l["value"], "ms", l["ms_per_step"], "e2e", l["e2e"]["value"], l["e2e"]["hostlink_frac"], "train", l["e2e_train"]["value"], l["e2e_train"]["hostlink_frac"], "ragged", {k: v["value"] for k, v in l["ragged"].items() if isinstance(v, dict)})
    except Exception as e:
        print(n, "FAILED", e)
PY
