from .image_resnet import ResNet
from .resnet import CustomResNet
from .unet import UNet

__all__ = ['CustomResNet', 'UNet']
