INFO = {
    "octmnist": {
        "python_class": "OCTMNIST",
        "task": "multi-class",
        "n_channels": 1,
        "label": {"0": "choroidal neovascularization", "1": "diabetic macular edema", "2": "drusen", "3": "normal"},
        "n_samples": {"train": 97477, "val": 10832, "test": 1000},
    }
}
