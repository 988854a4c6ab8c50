#!/bin/bash
# B200 drop-in for the reference's script/train/llava15_train_lora.sh: same flags, torchrun instead of the
# deepspeed launcher (one process per GPU, NCCL over NVLink; ZeRO-2 is implemented natively).
export PYTHONPATH=$PYTHONPATH:`realpath .`

task_name=llava15_7b_DPO
exp_name=llava15_rlaifv_lora
NGPU=${NGPU:-8}

python -m torch.distributed.run --nnodes=1 --nproc-per-node $NGPU --master-addr 127.0.0.1 --master-port ${MASTER_PORT:-29500} \
    -m rlaifv_b200.train_llava15_lora \
    --deepspeed ./script/zero2.json  \
    --model_name_or_path liuhaotian/llava-v1.5-7b \
    --data_dir ./RLAIF-V-Dataset_logps/ \
    --image_folder not_used \
    --vision_tower openai/clip-vit-large-patch14-336 \
    --mm_use_im_start_end False \
    --mm_use_im_patch_token False \
    --fully_tune False \
    --image_aspect_ratio pad \
    --bf16 True \
    --mm_projector_type mlp2x_gelu \
    --mm_vision_select_layer -2 \
    --output_dir .ckpt/$task_name-$exp_name/checkpoints \
    --num_train_epochs 10 \
    --per_device_train_batch_size 1 \
    --per_device_eval_batch_size 4 \
    --gradient_accumulation_steps 1 \
    --evaluation_strategy "no" \
    --save_strategy "steps" \
    --save_steps 167 \
    --save_total_limit 50 \
    --data_source_names '' \
    --data_source_weights 1 \
    --max_steps 2672 \
    --learning_rate 1e-5 \
    --weight_decay 0.01 \
    --warmup_ratio 0.05 \
    --lr_scheduler_type "cosine" \
    --logging_steps 2 \
    --logging_dir .ckpt/$task_name-$exp_name/log \
    --tf32 True \
    --model_max_length 2048 \
    --gradient_checkpointing True \
    --lazy_preprocess True \
    --task DPO \
    --report_to wandb \
    --run_name $exp_name \
    --dataloader_num_workers 16 \
    --dpo_use_average False \
    --dpo_token_weighted False \
    --dpo_token_weight 1.0 \
    --dpo_beta 0.1 \
    --lora_enable True

# same post-processing as the reference script: every checkpoint dir becomes loadable by llava/model/builder.py
parent_dir=".ckpt/$task_name-$exp_name/checkpoints"
for dir in "$parent_dir"/checkpoint-*; do
    if [ -d "$dir" ]; then
        new_dir="$parent_dir/RLAIFV7B-lora_$(basename "$dir")"
        mv "$dir" "$new_dir"
        cp "$parent_dir/config.json" "$parent_dir/non_lora_trainables.bin" "$new_dir/"
    fi
done
