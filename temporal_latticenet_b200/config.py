"""hjson cfg access with the reference's cfgParser interface (cfgParser.py:3-74): only the getters
the hot path consults."""
import hjson


class ConfigParser:
    def __init__(self, cfg_file):
        with open(cfg_file, "r") as f:
            self.contents = hjson.loads(f.read())

    def get_model_vars(self): return self.contents["model"]
    def get_lattice_gpu_vars(self): return self.contents["lattice_gpu"]
    def get_train_vars(self): return self.contents["train"]
    def get_eval_vars(self): return self.contents["eval"]

    def get_loader_vars(self):
        for mode in ("train", "eval"):
            name = self.contents.get(mode, {}).get("dataset_name")
            if name == "semantickitti":
                return self.contents["loader_semantic_kitti"]
            if name == "parislille":
                return self.contents["loader_paris_lille"]
        return None
