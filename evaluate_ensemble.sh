#!/bin/bash
# Positional-argument wrapper with the reference's argument order (evaluate_ensemble.sh of the
# reference); plot operations are outside the accelerated path and only print a notice.
# Set NGPUS>1 to shard the clips over several GPUs of this node.

PY="python"
if [ "${NGPUS:-1}" -gt 1 ]; then
    PY="python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NGPUS} --master-addr 127.0.0.1 --master-port ${MASTER_PORT:-29511}"
fi
COMMON=(-rf "Results" -af 3 -tmf "Trained_models/" -cs "unbalanced")

case "$1" in
  Global_evaluate_models|Combine_ensembles)
    # $2 = model list (word-split on purpose, like the reference), $3 = folds
    $PY evaluate_ensemble.py -op "$1" -mlist $2 -fn "$3" -is "test" "${COMMON[@]}" ;;
  Confusion_matrices|Difference_matrices)
    if [ "$2" = "Global" ]; then
      $PY evaluate_ensemble.py -op "$1" -et "$2" -mlist $3 -fn "$4" -is "test" "${COMMON[@]}"
    else
      $PY evaluate_ensemble.py -op "$1" -et "$2" -mt "$3" -tc "$4" -wt "$5" -b "$6" -w "$7" -ofs "$8" -as "$9" \
          -fn "${10}" -is "test" -hf_vei "Data/Weights/" "${COMMON[@]}"
    fi ;;
  *)
    $PY evaluate_ensemble.py -op "$1" -mt "$2" -tc "$3" -wt "$4" -b "$5" -w "$6" -ofs "$7" -as "$8" -fn "$9" \
        -is "${10}" -hf_vei "Data/Weights/" "${COMMON[@]}" ;;
esac
