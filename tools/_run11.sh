set -x
cd $GRAFT_REPO_ROOT
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_b11.json 2> gpurun_out/r2_b11.err; echo "bench rc=$?" >> gpurun_out/r2_b11.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_b11_ref.json 2> gpurun_out/r2_b11_ref.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke11.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2_smoke11.log
tail -3 gpurun_out/r2_b11.err; tail -4 gpurun_out/r2_smoke11.log
