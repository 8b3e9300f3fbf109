"""The reference's docs and launch scripts name this entry point but the fork does not ship it:
``python train_OpenAICLIP_stage2_only.py --config <yaml>`` (one process per GPU; torchrun for data parallelism).
The loop, checkpoint layout and YAML schema live in genhancer_b200/trainer.py."""
from genhancer_b200.trainer import main

if __name__ == "__main__":
    main("OpenAICLIP", "image", "stage2_only")
