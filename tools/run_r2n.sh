#!/bin/bash
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/r2n_pytest.log 2>&1; echo "tests exit $?"; tail -4 $O/r2n_pytest.log
B="timeout 600 python bench.py --mode train --steps 40 --no-cpu-baseline --no-parity"
for i in 1 2; do
$B > $O/r2n_train_early_pdl$i.json 2> $O/r2n_train_early_pdl$i.err; echo "early+pdl$i $?"; head -c 130 $O/r2n_train_early_pdl$i.json; echo
VP3D_PDL=0 $B > $O/r2n_train_early$i.json 2> $O/r2n_train_early$i.err; echo "early$i $?"; head -c 130 $O/r2n_train_early$i.json; echo
VP3D_PDL=0 $B --no-update-in-backward > $O/r2n_train_base$i.json 2> $O/r2n_train_base$i.err; echo "base$i $?"; head -c 130 $O/r2n_train_base$i.json; echo
$B --no-update-in-backward > $O/r2n_train_pdl$i.json 2> $O/r2n_train_pdl$i.err; echo "pdl$i $?"; head -c 130 $O/r2n_train_pdl$i.json; echo
VP3D_WGRAD_ORDER=before $B > $O/r2n_train_orderbefore$i.json 2> $O/r2n_train_orderbefore$i.err; echo "order-before$i $?"; head -c 130 $O/r2n_train_orderbefore$i.json; echo
VP3D_K1_2CTA=force $B > $O/r2n_train_force$i.json 2> $O/r2n_train_force$i.err; echo "force$i $?"; head -c 130 $O/r2n_train_force$i.json; echo
done
timeout 300 python bench.py --mode infer --steps 20 --no-cpu-baseline --no-parity > $O/r2n_infer_pdl.json 2> $O/r2n_infer_pdl.err; echo "infer pdl $?"; head -c 130 $O/r2n_infer_pdl.json; echo
VP3D_PDL=0 timeout 300 python bench.py --mode infer --steps 20 --no-cpu-baseline --no-parity > $O/r2n_infer_nopdl.json 2> $O/r2n_infer_nopdl.err; echo "infer nopdl $?"; head -c 130 $O/r2n_infer_nopdl.json; echo
