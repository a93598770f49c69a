"""CPU: the C-ABI shared library loads and exports every symbol include/colosseum_b200.h declares (no compute)."""
import ctypes
import os
import re

from conftest import ROOT
from colosseum_b200 import _cabi
from colosseum_b200.build import build_library


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "colosseum_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(colo_[a-z0-9_]+)\s*\(", txt)))


def test_library_builds_loads_and_exports_header():
    path = build_library()
    assert os.path.isfile(path)
    handle = ctypes.CDLL(path)
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(handle, n), f"{n} declared in the header but not exported"
    # and the ctypes prototypes cover exactly the header
    assert sorted(_cabi.PROTOTYPES) == names
    lib = _cabi.lib()
    assert lib.colo_version() >= 100
    assert lib.colo_launch_count() == 0


def test_struct_layouts_match_header():
    """field order of the ctypes mirrors == field order in the header structs"""
    txt = open(os.path.join(ROOT, "include", "colosseum_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    for cname, mirror in (("colo_backup_args", _cabi.BackupArgs), ("colo_mdp_tables", _cabi.MdpTables),
                          ("colo_env_batch", _cabi.EnvBatch), ("colo_resident_args", _cabi.ResidentArgs),
                          ("colo_env_server", _cabi.EnvServer), ("colo_qlearning_args", _cabi.QLearningArgs), ("colo_psrl_args", _cabi.PsrlArgs),
                          ("colo_ucrl2_args", _cabi.Ucrl2Args), ("colo_actor_args", _cabi.ActorArgs), ("colo_psrlc_args", _cabi.PsrlcArgs),
                          ("colo_suite_instance", _cabi.SuiteInstance), ("colo_suite_config", _cabi.SuiteConfig),
                          ("colo_suite_result", _cabi.SuiteResult)):
        end = txt.index("} " + cname + ";")
        body = txt[txt.rindex("typedef struct {", 0, end) + len("typedef struct {"):end]
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            for part in decl.split(","):
                part = re.sub(r"\[\d+\]", "", part)  # char error[160]
                fields.append(re.findall(r"[A-Za-z_][A-Za-z0-9_]*", part)[-1])
        assert fields == [f[0] for f in mirror._fields_], cname


def test_product_never_imports_oracle():
    """the oracle is test infrastructure: nothing under colosseum_b200/ may reference it"""
    pkg = os.path.join(ROOT, "colosseum_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "colo_oracle" not in src, f
