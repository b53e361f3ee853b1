"""Shadow of the reference's model_generator_progressive.py: same names, B200-native kernels."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import sisr_b200  # noqa: E402,F401
from sisr_b200.model_generator_progressive import (BasicBlock, GeneratorProgresiveBase,  # noqa: E402,F401
                                                    GeneratorSuffix)
