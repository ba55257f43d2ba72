"""TEST INFRASTRUCTURE — recipe for oracle/_ref: a runnable copy of the reference's own hot-path modules.

    python -m oracle.make_ref            (build container only; __graft_entry__.build() runs it when /root/reference exists)

Imports the reference's model files UNMODIFIED from /root/reference (through oracle/ref_loader.py + the monai 0.7.0
restatement in oracle/monai_compat), records which reference source files that import actually executed, and copies
exactly those files, byte for byte, to oracle/_ref/ (git-ignored: reference sources never enter the history; NOT
gpurun-ignored, so the directory travels to the GPU box like a built .so).  With oracle/_ref present,
`bench.py --impl reference` and `cpu_baseline` time the REFERENCE'S OWN nn.Modules (kind "reference") instead of the
oracle port.  monai itself remains the restatement (un-vendored dependency, absent offline).
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
DST = os.path.join(HERE, "_ref")


def main(quiet=False):
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    from oracle import ref_loader
    src_root = "/root/reference"
    if not os.path.isdir(os.path.join(src_root, "DosePrediction")):
        if not quiet:
            print("make_ref: /root/reference not present (GPU box): keeping the prebuilt oracle/_ref")
        return False
    os.environ["DP_REFERENCE_ROOT"] = src_root
    ref_loader.REFERENCE_ROOT = src_root
    ref_loader.dose_pyfer()
    ref_loader.oar_transeg()
    ref_loader.oar_transeg_old()
    ref_loader.loss()
    files = sorted({os.path.realpath(m.__file__) for m in list(sys.modules.values())
                    if getattr(m, "__file__", None) and os.path.realpath(m.__file__).startswith(src_root + os.sep)})
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    for f in files:
        rel = os.path.relpath(f, src_root)
        out = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(out), exist_ok=True)
        shutil.copyfile(f, out)
    with open(os.path.join(DST, "MANIFEST.txt"), "w") as fh:
        fh.write("# files copied unmodified from /root/reference by oracle/make_ref.py (git-ignored)\n")
        fh.writelines(os.path.relpath(f, src_root) + "\n" for f in files)
    if not quiet:
        print(f"make_ref: {len(files)} reference files -> {DST}")
    return True


if __name__ == "__main__":
    main()
