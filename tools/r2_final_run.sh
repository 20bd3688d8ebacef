# end-of-round verification on one B200: GPU parity suite, smoke, both bench arms (what the driver runs)
mkdir -p gpurun_out
(timeout 400 python -m pytest tests -m gpu -q > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log)
(timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2g_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2g_smoke.log)
(timeout 300 python bench.py --impl reference --steps 10 --warmup 1 > gpurun_out/r2g_bench_ref.json 2> gpurun_out/r2g_bench_ref.err; echo "ref rc=$?" >> gpurun_out/r2g_bench_ref.err)
(timeout 400 python bench.py > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; echo "bench rc=$?" >> gpurun_out/r2g_bench.err)
tail -3 gpurun_out/r2g_pytest.log; tail -2 gpurun_out/r2g_smoke.log; tail -2 gpurun_out/r2g_bench_ref.err; tail -2 gpurun_out/r2g_bench.err
