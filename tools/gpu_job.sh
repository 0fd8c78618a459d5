cd /root/repo
timeout 900 python -m pytest tests/test_vi.py tests/test_cli_surface.py -q > gpurun_out/r2_vitests.log 2>&1; echo "vi tests rc $?"
tail -15 gpurun_out/r2_vitests.log
timeout 300 python vi.py --env Tiger --solver ValueIteration --planning_horizon 6 --n_epochs 50 --max_steps 20 --seed 123 2>&1 | tail -12
