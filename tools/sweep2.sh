for mb in 2 3 4; do echo "MIN_BLOCKS=$mb"; PCR_MIN_BLOCKS=$mb bash tools/sweep.sh; done
