#!/bin/bash
# Round-2 multi-GPU evidence on ONE 8-GPU box (gpurun --gpus 8 -- bash scripts/multi_gpu_round.sh [tag]):
#   host-link ceiling at N=1/2/4/8, bench.py at N=8, BASELINE configs[4] (10k-clip corpus -> ViT batches) and configs[3]
#   (CNN batches of 128 into a ResNet18 six-head consumer) sharded over 8 ranks, and the serial-sum check of the gathered stats.
TAG=${1:-r02c}
OUT=gpurun_out
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi topo -m > $OUT/${TAG}_topo.txt 2>&1
rm -f $OUT/${TAG}_hostlink.jsonl
timeout 120 python scripts/hostlink_bench.py --out $OUT/${TAG}_hostlink.jsonl > /dev/null 2> $OUT/${TAG}_hostlink_n1.err
for n in 2 4 8; do
  timeout 180 $TR --nproc-per-node $n --master-port $((29600 + n)) scripts/hostlink_bench.py --out $OUT/${TAG}_hostlink.jsonl > /dev/null 2> $OUT/${TAG}_hostlink_n$n.err
done
echo "== hostlink done"; grep -c case $OUT/${TAG}_hostlink.jsonl
timeout 300 $TR --nproc-per-node 8 --master-port 29700 bench.py --gpus 8 --steps 10 --warmup 3 > $OUT/${TAG}_bench_n8.log 2> $OUT/${TAG}_bench_n8.err
echo "== bench n8 rc=$?"
timeout 300 $TR --nproc-per-node 8 --master-port 29701 scripts/corpus_stream.py --clips 10000 --mode vit --batch 50 --consume > $OUT/${TAG}_corpus_vit_10k_n8.log 2> $OUT/${TAG}_corpus_vit_10k_n8.err
echo "== corpus vit 10k n8 rc=$?"; tail -1 $OUT/${TAG}_corpus_vit_10k_n8.log
timeout 300 $TR --nproc-per-node 8 --master-port 29702 scripts/corpus_stream.py --clips 800 --mode cnn --batch 128 --model resnet18 > $OUT/${TAG}_corpus_cnn_resnet18_n8.log 2> $OUT/${TAG}_corpus_cnn_resnet18_n8.err
echo "== corpus cnn resnet18 n8 rc=$?"; tail -1 $OUT/${TAG}_corpus_cnn_resnet18_n8.log
timeout 300 $TR --nproc-per-node 8 --master-port 29703 scripts/corpus_stream.py --clips 800 --mode vit --batch 50 --consume > $OUT/${TAG}_corpus_vit_800_n8.log 2> $OUT/${TAG}_corpus_vit_800_n8.err
timeout 300 python scripts/corpus_stream.py --clips 800 --mode vit --batch 50 --consume > $OUT/${TAG}_corpus_vit_800_n1.log 2> $OUT/${TAG}_corpus_vit_800_n1.err
echo "== serial-sum check"; tail -1 $OUT/${TAG}_corpus_vit_800_n8.log | cut -c1-400; tail -1 $OUT/${TAG}_corpus_vit_800_n1.log | cut -c1-400
