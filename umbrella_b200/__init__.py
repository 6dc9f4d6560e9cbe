"""umbrella_b200 -- B200-native speculative-decoding verify path behind the UMbreLLa engine surface."""
__version__ = "0.1.0"
