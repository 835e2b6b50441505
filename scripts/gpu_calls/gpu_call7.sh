for a in "64" "128" "128 tornado" "128 curl box" "128 curl triangle 64" "128 curl triangle 256 0.002" "128 abc_flow"; do python scripts/diag_licvol.py $a 2>&1 | grep -v "^Volume"; done
