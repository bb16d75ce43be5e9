"""Build the approximate sweep at PCD_APX_LEVEL 1..4 (development tool) -> tools/libpcdist_apx<L>.so"""
import importlib.util, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("_pcd_build", os.path.join(ROOT, "3dpointcloudattack_b200", "build.py"))
build = importlib.util.module_from_spec(spec); spec.loader.exec_module(build)
for lvl in sys.argv[1:]:
    build.build(force=True, extra_flags=[f"-DPCD_APX_LEVEL={lvl}"], out=os.path.join(ROOT, "tools", f"libpcdist_apx{lvl}.so"))
    print("built level", lvl)
