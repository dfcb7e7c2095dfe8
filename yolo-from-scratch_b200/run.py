"""Launcher: `python -m yolo_from_scratch_b200.run /path/to/train.py [train.py arguments...]`.

Runs the reference CLI (train.py:1354-1564) unchanged, with the B200 path swapped in: the file is
imported as module `train` (so the import hook patches it) and then the body of its
`if __name__ == "__main__":` block is executed inside that module's namespace.
"""
import ast
import importlib
import os
import sys


def main_block(source, filename):
    tree = ast.parse(source, filename)
    body = []
    for node in tree.body:
        if (isinstance(node, ast.If) and isinstance(node.test, ast.Compare)
                and isinstance(node.test.left, ast.Name) and node.test.left.id == "__name__"):
            body.extend(node.body)
    return ast.Module(body=body, type_ignores=[])


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        raise SystemExit("usage: python -m yolo_from_scratch_b200.run /path/to/train.py [args...]")
    script = os.path.abspath(argv[0])
    from .install import enable_import_hook, install
    enable_import_hook()
    sys.path.insert(0, os.path.dirname(script))
    name = os.path.splitext(os.path.basename(script))[0]
    module = importlib.import_module(name)
    install(module)
    sys.argv = [script] + argv[1:]
    with open(script, encoding="utf-8") as f:
        code = compile(main_block(f.read(), script), script, "exec")
    exec(code, module.__dict__)


if __name__ == "__main__":
    main()
