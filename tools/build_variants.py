"""Build kernel-tuning variants of the library (select one with DW_LIB=<path>): python tools/build_variants.py name:DEF1,DEF2 ..."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from therldaisyworld_b200 import build as b
for spec in sys.argv[1:]:
    name, _, defs = spec.partition(":")
    print(b.build_variant(name, [d for d in defs.split(",") if d]))
