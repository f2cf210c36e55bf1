#!/bin/bash
# prints ms_total / ms_em / ms_genotype of the last of N resident scans: scripts/ms.sh [one_scan.py args]
python scripts/one_scan.py "$@" | tail -1 | python -c "import sys,ast; t=sys.stdin.read(); d=ast.literal_eval(t[:t.rindex('}')+1]); print('ms_total %.3f ms_em %.3f ms_genotype %.3f ms_screen %.3f' % (d['ms_total'], d['ms_em'], d['ms_genotype'], d['ms_screen']))"
