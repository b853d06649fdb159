from .base import AdaptorOutput, BaseAdaptor, BaseAdaptorConfig
from .general import OFAAdaptorConfig, OFAGeneralAdaptor, default_adaptor

__all__ = ["AdaptorOutput", "BaseAdaptor", "BaseAdaptorConfig", "OFAGeneralAdaptor", "default_adaptor", "OFAAdaptorConfig"]
