set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 8 4 2; do
  $TR --nproc-per-node $n --master-port 2951$n bench.py --gpus $n --steps 40 --warmup 3 > gpurun_out/mg_bench_n$n.log 2> gpurun_out/mg_bench_n$n.err
  $TR --nproc-per-node $n --master-port 2952$n tools/tilesplit.py grid 3840 2160 --frames 5 2>/dev/null | tail -1 > gpurun_out/mg_tile_grid4k_n$n.log
done
python bench.py --steps 40 --warmup 3 --no-cpu-baseline --no-deep-scene > gpurun_out/mg_bench_n1.log 2>/dev/null
python tools/tilesplit.py grid 3840 2160 --frames 5 2>/dev/null | tail -1 > gpurun_out/mg_tile_grid4k_n1.log
$TR --nproc-per-node 8 --master-port 29531 tools/flythrough.py --frames 600 2>/dev/null | tail -1 > gpurun_out/mg_fly_n8.log
$TR --nproc-per-node 8 --master-port 29532 tools/tilesplit.py soup 1920 1080 --tris 2000000 --frames 3 2>gpurun_out/mg_tile_soup_n8.err | tail -1 > gpurun_out/mg_tile_soup_n8.log
python tools/tilesplit.py soup 1920 1080 --tris 2000000 --frames 3 2>/dev/null | tail -1 > gpurun_out/mg_tile_soup_n1.log
python tools/flythrough.py --frames 600 2>/dev/null | tail -1 > gpurun_out/mg_fly_n1.log
tail -n 2 gpurun_out/mg_*.log | cut -c1-420
