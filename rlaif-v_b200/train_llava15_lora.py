"""LoRA-DPO entry point — drop-in for `muffin/train/train_llava15_lora.py` (flags of
script/train/llava15_train_lora.sh:6-49; LoRA arguments `:112-117`: r=64, alpha=16, dropout=0.05, bias none).

The adapter path lives in the same engine as full fine-tuning (model.LoraStore); this module only
makes `--lora_enable` default to what the reference's LoRA script passes and keeps the module name the script calls.
"""
import sys

from .train_llava15 import train as _train


def train(attn_implementation=None, argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if "--lora_enable" not in argv:
        argv += ["--lora_enable", "True"]
    return _train(attn_implementation=attn_implementation, argv=argv)


if __name__ == "__main__":
    train(attn_implementation="flash_attention_2")
