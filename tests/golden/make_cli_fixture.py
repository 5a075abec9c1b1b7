"""Generates tests/golden/reference_cli_options.json from the reference's click options (mepp/cli.py:12-456): for every
option its flags, destination name, default (as source text) and whether it is a flag / required.  The file is parsed
as text -- importing it needs TensorFlow, matplotlib etc.  Run in the build container (needs /root/reference)."""
import ast
import json
import os

SRC = "/root/reference/mepp/cli.py"
tree = ast.parse(open(SRC).read())
main = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "main")
options = []
for dec in main.decorator_list:
    if not (isinstance(dec, ast.Call) and getattr(dec.func, "attr", "") == "option"):
        continue
    strs = [a.value for a in dec.args if isinstance(a, ast.Constant) and isinstance(a.value, str)]
    flags = [s for s in strs if s.startswith("-")]
    dest = next((s for s in strs if not s.startswith("-")), None)
    kw = {k.arg: ast.unparse(k.value) for k in dec.keywords}
    options.append(dict(flags=flags, dest=dest, default=kw.get("default"), is_flag=kw.get("is_flag") == "True",
                        flag_value=kw.get("flag_value"), required=kw.get("required") == "True",
                        nargs=kw.get("nargs"), type=kw.get("type")))
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_cli_options.json")
with open(out, "w") as f:
    json.dump(options, f, indent=1)
print(len(options), "options ->", out)
