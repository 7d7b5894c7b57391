"""Worker-side public API of elasticai_api (reference: /root/reference/elasticai_api/),
re-hosted on torch.distributed: NCCL over NVLink/NVSwitch on GPUs (gloo for CPU tests)
instead of Horovod+Gloo (elasticai_api/common/base_controller.py:100-101)."""
