from .awq_utils import AwqLinear  # noqa: F401
