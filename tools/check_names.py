"""Poor man's pyflakes (none in this image): names a scope reads as globals that the module never binds."""
import builtins
import symtable
import sys


def check(path):
    src = open(path).read()
    top = symtable.symtable(src, path, "exec")
    module_names = {s.get_name() for s in top.get_symbols() if s.is_assigned() or s.is_imported() or s.is_namespace()}
    known = module_names | set(dir(builtins)) | {"__file__", "__name__", "__doc__"}
    bad = []

    def walk(tab):
        for s in tab.get_symbols():
            if s.is_referenced() and s.is_global() and s.get_name() not in known:
                bad.append((tab.get_lineno(), tab.get_name(), s.get_name()))
        for ch in tab.get_children():
            walk(ch)
    walk(top)
    return bad


if __name__ == "__main__":
    rc = 0
    for p in sys.argv[1:]:
        for line, scope, name in check(p):
            print(f"{p}:{line}: in {scope}: undefined name {name}")
            rc = 1
    sys.exit(rc)
