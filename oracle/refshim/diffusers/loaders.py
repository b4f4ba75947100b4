class FromOriginalControlnetMixin:
    pass


class UNet2DConditionLoadersMixin:
    pass
