# End-of-round verification recipe (what the numbers in profiles/r02_* were taken with).  Run each block under gpurun with the GPU count it names.
#   1 GPU :  bash tools/verify_round.sh n1
#   2 GPUs:  bash tools/verify_round.sh n2     (also runs the 2-GPU byte-identity tests a 1-GPU box skips)
#   8 GPUs:  bash tools/verify_round.sh n8     (N = 8, then N = 4 on the same box)
mkdir -p gpurun_out
case "$1" in
n1)
    python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/verify_tests.log; tail -2 gpurun_out/verify_tests.log
    python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
    python bench.py --steps 20 --warmup 5 > gpurun_out/verify_bench_n1.json 2> gpurun_out/verify_bench_n1.err; echo rc=$?
    ;;
n2)
    python -m pytest tests/test_gpu_distributed.py tests/test_gpu_kzg.py -x -q -m gpu 2>&1 | tail -4 > gpurun_out/verify_tests_2gpu.log; cat gpurun_out/verify_tests_2gpu.log
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline \
        > gpurun_out/verify_bench_n2.json 2> gpurun_out/verify_bench_n2.err; echo rc2=$?
    ;;
n8)
    for n in 8 4; do
        python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29510 + n)) bench.py --gpus $n --steps 20 --warmup 5 --no-cpu-baseline \
            > gpurun_out/verify_bench_n$n.json 2> gpurun_out/verify_bench_n$n.err; echo rc$n=$?
    done
    ;;
*) echo "usage: bash tools/verify_round.sh n1|n2|n8"; exit 2;;
esac
python - <<'P'
import glob, json
for f in sorted(glob.glob('gpurun_out/verify_bench_n*.json')):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d['n_gpus'], 'value', round(d['value'], 3), 'e2e', round(d['e2e']['value'], 3), {k: round(v.get('ms', 0), 3) for k, v in d.get('configs', {}).items()})
    except Exception as e:
        print(f, 'unreadable:', e)
P
