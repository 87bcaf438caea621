# Same-box A/B of programmatic dependent launch on the graphed train step (WSAE_PDL=1 opts in).
set -x
for i in 1 2 3; do
for p in 0 1; do
echo "pdl=$p"; WSAE_PDL=$p python bench.py --steps 50 --warmup 10 --value-only 2>&1 | grep -o '"ms_per_step": [0-9.]*'
done; done
for p in 0 1 0 1; do
echo "small pdl=$p"; WSAE_PDL=$p python bench.py --steps 200 --warmup 20 --batch 128 --value-only 2>&1 | grep -o '"ms_per_step": [0-9.]*'
done
