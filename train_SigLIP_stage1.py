"""Same entry point as /root/reference/Continuous/train_SigLIP_stage1.py:
``python train_SigLIP_stage1.py --config <yaml>`` (one process per GPU; torchrun for data parallelism).
The loop, checkpoint layout and YAML schema live in genhancer_b200/trainer.py."""
from genhancer_b200.trainer import main

if __name__ == "__main__":
    main("SigLIP", "image", "stage1")
