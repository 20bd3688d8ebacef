# round-2 late verification: GPU parity suite, smoke, bench (own arm + reference arm)
mkdir -p gpurun_out
(timeout 300 python -m pytest tests -m gpu -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log)
(timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2d_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2d_smoke.log)
(timeout 400 python bench.py > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo "bench rc=$?" >> gpurun_out/r2d_bench.err)
tail -3 gpurun_out/r2d_pytest.log; tail -2 gpurun_out/r2d_smoke.log; tail -2 gpurun_out/r2d_bench.err
