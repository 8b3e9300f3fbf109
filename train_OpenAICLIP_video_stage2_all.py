"""Same entry point as /root/reference/Continuous/train_OpenAICLIP_video_stage2_all.py:
``python train_OpenAICLIP_video_stage2_all.py --config <yaml>`` (one process per GPU; torchrun for data parallelism).
The loop, checkpoint layout and YAML schema live in genhancer_b200/trainer.py."""
from genhancer_b200.trainer import main

if __name__ == "__main__":
    main("OpenAICLIP", "video", "stage2_all")
