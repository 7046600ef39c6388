"""Run the reference's own `test_rtf.py` unmodified against a golf_b200 YAML.

    python tools/run_test_rtf.py <config.yaml> <ckpt> <wav> [--cuda] [-n 10] [--duration 6]

Needs the reference tree (GOLF_REFERENCE_ROOT, default /root/reference) and -- for `--cuda` -- a GPU.  What it adds to
`python /root/reference/test_rtf.py ...`: the absent third-party modules (lightning, torchlpc, kazane, ...) are provided
by the stand-ins of oracle/refimport.py / oracle/lightning_standin.py, and with `--b200` the decoder's class paths in
the config are rewritten from `models.*` to `golf_b200.*` on the fly so that a shipped
ckpts/*/config.yaml can be passed as it is."""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import yaml

    from oracle import refimport

    argv = sys.argv[1:]
    if "--b200" in argv:
        argv.remove("--b200")
        cfg = yaml.safe_load(open(argv[0]))
        model = cfg["model"].get("init_args", cfg["model"])

        def rewrite(c):
            if isinstance(c, dict):
                return {k: (v.replace("models.", "golf_b200.", 1) if k == "class_path" else rewrite(v)) for k, v in c.items()}
            return c

        model["decoder"] = rewrite(model["decoder"])
        tmp = tempfile.NamedTemporaryFile("w", suffix=".yaml", delete=False)
        yaml.safe_dump(cfg, tmp)
        tmp.close()
        argv[0] = tmp.name
    _, rtf = refimport.import_harness()
    import torchaudio

    try:
        import torchcodec  # noqa: F401  (torchaudio 2.11 loads audio through it)
    except ImportError:  # absent from this image: read PCM / float wav files with scipy instead

        def _load(path, *a, **k):
            import numpy as np
            import torch
            from scipy.io import wavfile

            sr, x = wavfile.read(path)
            x = x.astype(np.float32) if x.dtype.kind == "f" else x.astype(np.float32) / float(np.iinfo(x.dtype).max + 1)
            return torch.from_numpy(x.reshape(len(x), -1).T.copy()), sr

        torchaudio.load = _load
    sys.argv = [os.path.join(refimport.REF_ROOT, "test_rtf.py")] + argv
    rtf.main()


if __name__ == "__main__":
    main()
