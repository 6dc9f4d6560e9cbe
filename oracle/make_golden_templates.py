"""TEST INFRASTRUCTURE.  Dumps the reference's chat-template tables (umbrella/templates.py:1-47) into
tests/golden/templates.json so the product's assembled tables can be pinned byte for byte on a box without
/root/reference.      python oracle/make_golden_templates.py
"""
import importlib.util
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("ref_templates", "/root/reference/umbrella/templates.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)
out = {"Prompts": ref.Prompts, "SysPrompts": ref.SysPrompts, "ExtraPrompts": ref.ExtraPrompts}
with open(os.path.join(ROOT, "tests", "golden", "templates.json"), "w") as f:
    json.dump(out, f, indent=1, sort_keys=True)
print({k: sorted(v) for k, v in out.items()})
