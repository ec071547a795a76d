"""Configuration-driven access to the analysis products (drop-in for drift/core/manager.py).

``ProductManager.from_config(path)`` takes the reference's YAML layout: a ``config:`` section
(output directory, which stages to run, BeamTransfer options), a ``telescope:`` section whose
``type:`` is a registered name or a ``{class, module, file}`` mapping for user-defined telescopes,
and an optional ``kltransform:`` list.  It installs ``<output_directory>/config.yaml``, builds the
telescope, the ``BeamTransfer`` variant asked for and the KL transforms; ``generate()`` runs the
beam-transfer, SVD and KL stages on the GPU.  ``psfisher`` entries are accepted and skipped with a
warning (power-spectrum estimation is outside this package).
"""

import importlib
import importlib.util
import logging
import os
import sys
import warnings

import yaml

from .. import parallel
from ..telescope import cylinder, exotic_cylinder, restrictedcylinder
from . import beamtransfer, doublekl, kltransform

logger = logging.getLogger(__name__)

kltype_dict = {"KLTransform": kltransform.KLTransform, "DoubleKL": doublekl.DoubleKL}

# manager.py:28-38 of the reference.  Not built: GMRT (needs the antenna-position data file) and
# FocalPlane (the reference class itself cannot be instantiated: it lacks the abstract `beamclass`)
teltype_dict = {
    "UnpolarisedCylinder": cylinder.UnpolarisedCylinderTelescope,
    "PolarisedCylinder": cylinder.PolarisedCylinderTelescope,
    "RestrictedCylinder": restrictedcylinder.RestrictedCylinder,
    "RestrictedPolarisedCylinder": restrictedcylinder.RestrictedPolarisedCylinder,
    "RestrictedExtra": restrictedcylinder.RestrictedExtra,
    "GradientCylinder": exotic_cylinder.GradientCylinder,
    "PertCylinder": exotic_cylinder.CylinderPerturbed,
}


def _resolve_class(clstype, clsdict, objtype=""):
    """Registered name or ``{module, class[, file]}`` mapping -> class (manager.py:54-80)."""
    if isinstance(clstype, dict):
        modname, clsname = clstype["module"], clstype["class"]
        if "file" in clstype:
            spec = importlib.util.spec_from_file_location(modname, clstype["file"])
            module = importlib.util.module_from_spec(spec)
            # registered like imp.load_source does (manager.py:69 of the reference): pickling the
            # telescope (BeamTransfer.generate) looks the class up through sys.modules
            sys.modules[modname] = module
            spec.loader.exec_module(module)
        else:
            module = importlib.import_module(modname)
        return module.__dict__[clsname]
    if clstype in clsdict:
        return clsdict[clstype]
    raise Exception(f"Unsupported {objtype}")


class ProductManager(object):
    """Telescope + beam-transfer products of one configuration (manager.py:83-305)."""

    directory = None
    gen_beams = False
    gen_kl = False
    gen_ps = False
    gen_proj = False
    skip_svd = False
    skip_svd_inv = False

    @staticmethod
    def _installed_config(configfile):
        """Path of ``<output_directory>/config.yaml``, written from ``configfile`` (with the output
        directory made absolute) unless it is that very file (manager.py:119-166)."""
        comm = parallel.Comm.current()
        configfile = os.path.normpath(os.path.expandvars(os.path.expanduser(configfile)))
        if not os.path.exists(configfile):
            raise Exception(f"Configuration file does not exist {configfile}.")
        if os.path.isdir(configfile):  # a product directory stands for the copy inside it
            configfile = configfile + "/config.yaml"
        with open(configfile, "r") as f:
            given = yaml.safe_load(f)["config"]["output_directory"]
        outdir = given
        if not os.path.isabs(outdir):  # relative to the configuration file, not to the cwd
            outdir = os.path.abspath(os.path.normpath(os.path.join(os.path.dirname(configfile), outdir)))
        installed = os.path.join(outdir, "config.yaml")
        if comm.rank0:
            os.makedirs(outdir, exist_ok=True)
            if not os.path.exists(installed) or not os.path.samefile(configfile, installed):
                with open(configfile, "r") as f:
                    text = f.read()
                if given != outdir:
                    text = text.replace(given, outdir)
                with open(installed, "w+") as f:
                    f.write(text)
        comm.barrier()
        return installed

    @classmethod
    def from_config(cls, configfile):
        manager = cls()
        with open(cls._installed_config(configfile)) as f:
            manager.apply_config(yaml.safe_load(f))
        return manager

    def apply_config(self, yconf):
        if "config" not in yconf:
            raise ValueError("Configuration file must have an 'config' section.")
        if "telescope" not in yconf:
            raise ValueError("Configuration file must have an 'telescope' section.")
        self.config = yconf
        self.directory = os.path.expandvars(os.path.expanduser(yconf["config"]["output_directory"]))
        logger.info(f"Product directory: {self.directory}")

        telclass = _resolve_class(yconf["telescope"]["type"], teltype_dict, "telescope")
        self.telescope = telclass.from_config(yconf["telescope"])

        conf = yconf["config"]
        if conf.get("reionisation"):  # manager.py:210-211 of the reference
            from . import skymodel

            skymodel._reionisation = True
        # BeamTransfer variant (manager.py:217-221): plain, no SVD, or the single full SVD
        btclass = beamtransfer.BeamTransfer
        if conf.get("nosvd"):
            btclass = beamtransfer.BeamTransferNoSVD
        if conf.get("fullsvd"):
            btclass = beamtransfer.BeamTransferFullSVD
        self.beamtransfer = btclass(self.directory + "/bt/", telescope=self.telescope)
        self.beamtransfer.read_config(conf)

        self.gen_beams = bool(conf.get("beamtransfers"))
        self.skip_svd = bool(conf.get("skip_svd"))
        self.gen_kl = bool(conf.get("kltransform"))
        self.gen_ps = bool(conf.get("psfisher"))
        # KL transforms (manager.py:232-246)
        self.kltransforms = {}
        for klentry in yconf.get("kltransform", []) or []:
            klclass = _resolve_class(klentry["type"], kltype_dict, "KL filter")
            self.kltransforms[klentry["name"]] = klclass.from_config(klentry, self.beamtransfer,
                                                                      subdir=klentry["name"])
        self.psestimators = {}
        if "psfisher" in yconf:
            warnings.warn("`psfisher` entries are accepted but not generated by driftscan_b200")

    def generate(self):
        os.makedirs(self.directory, exist_ok=True)
        if parallel.Comm.current().rank0:
            with open(os.path.join(self.directory, "configdump.yaml"), "w") as fh:
                yaml.dump(self.config, fh)
        if self.gen_beams:
            self.beamtransfer.generate(skip_svd=self.skip_svd)
        if self.gen_kl:
            for klname, klobj in self.kltransforms.items():
                klobj.generate()
        if self.gen_ps:
            logger.warning("power-spectrum estimation is outside the scope of this build; skipped")
        logger.info("DONE GENERATING PRODUCTS")
