"""B200-native engine for the tiled U-Net prediction path of bio-image-unet (drop-in Python surface)."""
__version__ = '0.1.0'
