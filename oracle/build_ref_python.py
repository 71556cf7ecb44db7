"""TEST INFRASTRUCTURE: build the UNMODIFIED reference's Python binding (python/build_ceed_cffi.py, package `libceed`) into
oracle/_ref/python/, linked against oracle/_ref/lib-cuda/libceed.so.  The reference's own cffi build script is executed from where it lies
under /root/reference (it reads the public headers relative to the source root); only its link / include directories are redirected to the
oracle/_ref install, and the pure-Python modules of the package are staged next to the extension, as the reference's setup.py does
(package_dir={"libceed": "python"}).  Nothing is copied into the repository: oracle/_ref is git-ignored.
usage: python oracle/build_ref_python.py <reference root> <oracle/_ref>"""
import glob
import os
import runpy
import shutil
import sys

ref, out = os.path.abspath(sys.argv[1]), os.path.abspath(sys.argv[2])
pkg = os.path.join(out, "python")
os.makedirs(os.path.join(pkg, "libceed"), exist_ok=True)
os.chdir(ref)
ns = runpy.run_path(os.path.join(ref, "python", "build_ceed_cffi.py"), run_name="build_ceed_cffi")
ffibuilder = ns["ffibuilder"]
name, source, ext, kwds = ffibuilder._assigned_source
kwds["include_dirs"] = [os.path.join(out, "include")]
kwds["library_dirs"] = [os.path.join(out, "lib-cuda")]
kwds["runtime_library_dirs"] = ["$ORIGIN/../lib-cuda"]
kwds["extra_compile_args"] = ["-w", "-O1"]
ffibuilder._assigned_source = (name, source, ext, kwds)
ffibuilder.compile(tmpdir=pkg, verbose=False)
for f in glob.glob(os.path.join(ref, "python", "*.py")):
    if os.path.basename(f) != "build_ceed_cffi.py":
        shutil.copy(f, os.path.join(pkg, "libceed"))
for junk in glob.glob(os.path.join(pkg, "_ceed_cffi.c")) + glob.glob(os.path.join(pkg, "_ceed_cffi.o")):
    os.remove(junk)
print("built", glob.glob(os.path.join(pkg, "_ceed_cffi*.so")))
