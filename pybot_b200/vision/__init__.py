"""pybot.vision drop-in for the hot path: camera models, matcher, BA residuals."""
