set -x
O=gpurun_out
T=${1:-r3c}
shift
timeout 1200 python -W ignore::UserWarning -m pytest "$@" -m gpu -q -x > $O/${T}_pytest.log 2>&1; echo "rc=$?" >> $O/${T}_pytest.log
tail -n 25 $O/${T}_pytest.log
