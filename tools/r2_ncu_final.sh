#!/bin/bash
# final round-2 ncu evidence with the committed defaults: launch list + whole-trajectory captures (C2: 9 iterations = 18 launches)
LAUNCHES=1 WLS=c2 NCU_COUNT=18 timeout 600 bash tools/r2_ncu_lists.sh
LAUNCHES=0 WLS="c3 c4" NCU_COUNT=12 timeout 700 bash tools/r2_ncu_lists.sh
