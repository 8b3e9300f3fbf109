"""Same entry point as /root/reference/Continuous/train_MetaCLIP_stage2_only.py:
``python train_MetaCLIP_stage2_only.py --config <yaml>`` (one process per GPU; torchrun for data parallelism).
The loop, checkpoint layout and YAML schema live in genhancer_b200/trainer.py."""
from genhancer_b200.trainer import main

if __name__ == "__main__":
    main("MetaCLIP", "image", "stage2_only")
