for d in 0 14 15 1; do echo "impl 2 debug mask $d"; RSR_IMPL=2 RSR_DEBUG=$d RSR_TIME_CONVS=1 timeout 60 python tools/run_case.py 1920 1080 400 3 | tail -1; done
echo "impl 2 plain"; RSR_IMPL=2 timeout 60 python tools/run_case.py 1920 1080 400 4 | tail -2
echo "impl 1 plain"; RSR_IMPL=1 timeout 60 python tools/run_case.py 1920 1080 400 4 | tail -2
