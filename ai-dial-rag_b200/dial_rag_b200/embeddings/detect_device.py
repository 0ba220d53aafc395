"""Device selection, mirror of aidial_rag/embeddings/detect_device.py with one more value:
``b200`` (== ``cuda`` here: the only execution path of this package is the sm_100a library)."""

from __future__ import annotations

from enum import StrEnum


class DeviceType(StrEnum):
    AUTO = "auto"
    CPU = "cpu"
    CUDA = "cuda"
    B200 = "b200"


def autodetect_device() -> DeviceType:
    import torch

    return DeviceType.B200 if torch.cuda.is_available() else DeviceType.CPU


def detect_device(device_str: str) -> DeviceType:
    if device_str == DeviceType.AUTO:
        return autodetect_device()
    if device_str in list(DeviceType):
        return DeviceType(device_str)
    raise ValueError(f"Unknown device type: {device_str}")
