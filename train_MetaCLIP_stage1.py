"""Same entry point as /root/reference/Continuous/train_MetaCLIP_stage1.py:
``python train_MetaCLIP_stage1.py --config <yaml>`` (one process per GPU; torchrun for data parallelism).
The loop, checkpoint layout and YAML schema live in genhancer_b200/trainer.py."""
from genhancer_b200.trainer import main

if __name__ == "__main__":
    main("MetaCLIP", "image", "stage1")
