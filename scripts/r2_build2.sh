NBK_BUILD_TRACE=1 python scripts/build_profile.py 1024 3 2>&1 | grep -E "whole call|build_ms|arena allocated" | tail -9
python scripts/build_profile.py 512 4 | tail -4
python -m pytest tests -m gpu -x -q -k "build or tree or tiny or concurrent or multi or replica or empty or fixture or config1 or second_device" 2>&1 | tail -3
python scripts/config_sweep.py --configs 5 --out gpurun_out/r2_config5.json 2>&1 | cut -c1-200 | tail -2
