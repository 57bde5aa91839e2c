"""Identity-augmentation slice of the reference's evaluation tools (videoseal/evals/) running against this backend."""
