#!/bin/bash
# exercises the reference-style entry point end to end on synthetic data (tiny run)
cd food101-super-resolution_b200
WANDB_MODE=disabled SR_SYNTHETIC_DATA=96 SRK_DTYPE=bf16 timeout 300 python train.py --architecture RESNET --batch_size 16 --epochs 2 --loss_function nlpd --patience 3 --save_name smoke_resnet 2>&1 | tail -6
WANDB_MODE=disabled SR_SYNTHETIC_DATA=64 SRK_DTYPE=fp32 timeout 300 python train.py --architecture SRCNN --batch_size 16 --epochs 1 --loss_function mae --save_name smoke_srcnn 2>&1 | tail -3
rm -rf weights
