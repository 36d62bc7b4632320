# N-GPU e2e with the panel uploaded by every rank (0) or replicated over NVLink in 1 / 4 / 8 pieces.  Diagnostic.
N=${N:-2}
for p in ${PIECES:-0 1 4 8}; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+p)) \
    bench.py --gpus $N --no-cpu-baseline --panel-pieces $p > gpurun_out/n${N}_p$p.json 2> gpurun_out/n${N}_p$p.err
  echo "pieces=$p rc=$?"
done
