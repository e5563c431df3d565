cd /root/repo
python - "$@" <<'PY'
import importlib.util, os, sys
def build(suffix, flags):
    os.environ["HEXL_B200_BUILD_SUFFIX"]=suffix; os.environ["HEXL_B200_BUILD_FLAGS"]=flags
    spec = importlib.util.spec_from_file_location("_b"+suffix, "hexl_b200/build.py")
    m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m); print(m.build())
import sys
for a in sys.argv[1:]:
    sfx, fl = a.split("=",1)
    build(sfx, fl)
PY
