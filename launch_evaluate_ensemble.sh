#!/bin/bash
# Interactive front end with the reference's prompts (launch_evaluate_ensemble.sh of the reference).
# The reference submits through `sbatch`; here the job runs directly unless USE_SBATCH=1.

RUN="bash"
if [ "${USE_SBATCH:-0}" = "1" ]; then RUN="sbatch"; fi
DEFAULT_GLOBAL="SPECIALCASE_PRETRAINED R3D_34_SCRATCH TWOSTREAM_I3D_PRETRAINED TWOSTREAM_I3D_SCRATCH C3D_PRETRAINED C3D_SCRATCH I3D_PRETRAINED I3D_SCRATCH"

ask() { echo "$1"; read REPLY_VALUE; }

ask "Insert the operation name : ['Confusion_matrices', 'Difference_matrices', 'Evaluate_ensembles', 'Store_models_probabilities', 'StickDiagrams_wellClassifiedClips_per_numberOfModels', 'Global_evaluate_models', 'Combine_ensembles']"
operation=$REPLY_VALUE

if [ "$operation" = "Global_evaluate_models" ] || [ "$operation" = "Combine_ensembles" ]; then
    ask "Insert the number of folds"; folds_number=$REPLY_VALUE
    ask "Would like to mention the models to integrate in the global ensemble ? [Yes/No]"; integrate=$REPLY_VALUE
    ask "Which sets are invovled ? [test/train_val]"; involved_sets=$REPLY_VALUE
    models_list="$DEFAULT_GLOBAL"
    if [ "$integrate" = "Yes" ]; then
        ask "What is the list of models : Example TWOSTREAM_I3D_PRETRAINED"; models_list=$REPLY_VALUE
    fi
    $RUN evaluate_ensemble.sh "$operation" "$models_list" "$folds_number" "$involved_sets"
    exit $?
fi

if [ "$operation" = "Confusion_matrices" ] || [ "$operation" = "Difference_matrices" ]; then
    ask "Insert the ensemble type [Unique/Global]"; ensemble_type=$REPLY_VALUE
    if [ "$ensemble_type" = "Global" ]; then
        ask "Insert the number of folds"; folds_number=$REPLY_VALUE
        ask "What is the list of models : Example TWOSTREAM_I3D_PRETRAINED TWOSTREAM_I3D_SCRATCH"; models_list=$REPLY_VALUE
        $RUN evaluate_ensemble.sh "$operation" "$ensemble_type" "$models_list" "$folds_number"
        exit $?
    fi
fi

ask "Choose any of the following model types : [TWOSTREAM_I3D,I3D,C3D,R3D_18,R3D_34,R3D_50,R3D_101,R3D_152]"; model_type=$REPLY_VALUE
ask "Choose any of the training preconditions : [_PRETRAINED,_SCRATCH]"; training_condition=$REPLY_VALUE
ask "Insert the augmentation status : ['non_augmented', 'augmented_onTheFly', 'augmented_precomputed']"; augmentation_status=$REPLY_VALUE
ask "Insert the optical flow status : ['TVL1_precomputed', 'FarneBack_onTheFly']"; optical_flow_status=$REPLY_VALUE
ask "Insert the weighting method type : ['GRID_SEARCH', 'DIFFERENTIAL_EVOLUTION', 'SUM', 'VALIDATION_ERROR_INVERSE', 'MAXIMUM']"; weights_type=$REPLY_VALUE
ask "Insert batch_size"; batch_size=$REPLY_VALUE
ask "Insert the number of workers"; workers=$REPLY_VALUE
ask "Insert the number of folds"; folds_number=$REPLY_VALUE
if [ "$operation" = "Confusion_matrices" ] || [ "$operation" = "Difference_matrices" ]; then
    $RUN evaluate_ensemble.sh "$operation" "$ensemble_type" "$model_type" "$training_condition" "$weights_type" "$batch_size" "$workers" "$optical_flow_status" "$augmentation_status" "$folds_number"
else
    ask "Which sets are invovled ? [test/train_val]"; involved_sets=$REPLY_VALUE
    $RUN evaluate_ensemble.sh "$operation" "$model_type" "$training_condition" "$weights_type" "$batch_size" "$workers" "$optical_flow_status" "$augmentation_status" "$folds_number" "$involved_sets"
fi
