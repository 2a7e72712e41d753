"""Default settings of the command-line tool and the model classes.

Same keys and values as the reference's ``scvae/defaults.json`` (loaded there through
``scvae/defaults.py:21-24``) so that ``None`` arguments resolve identically.
"""

defaults = {
    "data": {
        "format": "infer", "directory": "data", "map_features": False,
        "feature_selection": [], "example_filter": [], "preprocessing_methods": [],
        "noisy_preprocessing_methods": [], "split_data_set": False,
        "splitting_method": "default", "splitting_fraction": 0.9,
    },
    "analyses": {
        "directory": "analyses", "decomposition_method": "PCA",
        "decomposition_dimensionality": 2, "highlight_feature_indices": [],
        "included_analyses": "standard", "analysis_level": "normal", "export_options": [],
    },
    "models": {
        "directory": "models", "type": "VAE", "latent_size": 2, "hidden_sizes": [100],
        "number_of_samples": {"training": 1, "evaluation": 1},
        "latent_distribution": {"VAE": "gaussian", "GMVAE": "gaussian mixture"},
        "number_of_classes": 1, "parameterise_latent_posterior": False,
        "inference_architecture": "MLP", "generative_architecture": "MLP",
        "reconstruction_distribution": "poisson", "number_of_reconstruction_classes": 0,
        "prior_probabilities_method": "uniform", "number_of_warm_up_epochs": 0,
        "kl_weight": 1, "proportion_of_free_nats_for_y_kl_divergence": 0.0,
        "minibatch_normalisation": True, "batch_correction": False,
        "dropout_keep_probabilities": [], "count_sum": False, "number_of_epochs": 200,
        "minibatch_size": 100, "learning_rate": 1e-4, "sample_size": 0, "run_id": "",
        "new_run": False, "reset_training": False,
    },
    "evaluation": {
        "data_set_kind": "test", "prediction_training_set_kind": "training",
        "prediction_method": "", "model_versions": "all",
    },
    "cross_analysis": {"log_summary": False},
}
