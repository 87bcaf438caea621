set -x
O=gpurun_out; T=r4m
python tools/profile_small_batch.py 128 > $O/${T}_base.txt 2>&1
WSAE_WGRAD=scatter python tools/profile_small_batch.py 128 > $O/${T}_scatter.txt 2>&1
cat $O/${T}_base.txt $O/${T}_scatter.txt
