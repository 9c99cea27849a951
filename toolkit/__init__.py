"""Drop-in `toolkit` package: the reference's Python surface (toolkit.models.get_models, the model class,
toolkit.utils.loss) re-exported from the B200-native implementation in sdumc_b200/."""
