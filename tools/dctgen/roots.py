import math
import numpy as np
f64=np.float64
class UnityRoots:
    def __init__(self,n):
        self.N=n
        pi=3.141592653589793238462643383279502884197
        ang=0.25*pi/n
        self.ang=ang
        nval=(n+2)//2
        shift=1
        while (1<<shift)*(1<<shift)<nval: shift+=1
        self.shift=shift; self.mask=(1<<shift)-1
        self.v1=[(1.0,0.0)]+[self.calc(i,n,ang) for i in range(1,self.mask+1)]
        n2=(nval+self.mask)//(self.mask+1)
        self.v2=[(1.0,0.0)]+[self.calc(i*(self.mask+1),n,ang) for i in range(1,n2)]
    @staticmethod
    def calc(x,n,ang):
        x<<=3
        c,s=math.cos,math.sin
        if x<4*n:
            if x<2*n:
                if x<n: return (c(float(x)*ang), s(float(x)*ang))
                return (s(float(2*n-x)*ang), c(float(2*n-x)*ang))
            else:
                x-=2*n
                if x<n: return (-s(float(x)*ang), c(float(x)*ang))
                return (-c(float(2*n-x)*ang), s(float(2*n-x)*ang))
        else:
            x=8*n-x
            if x<2*n:
                if x<n: return (c(float(x)*ang), -s(float(x)*ang))
                return (s(float(2*n-x)*ang), -c(float(2*n-x)*ang))
            else:
                x-=2*n
                if x<n: return (-s(float(x)*ang), -c(float(x)*ang))
                return (-c(float(2*n-x)*ang), -s(float(2*n-x)*ang))
    def __getitem__(self,idx):
        if 2*idx<=self.N:
            x1=self.v1[idx&self.mask]; x2=self.v2[idx>>self.shift]
            return (x1[0]*x2[0]-x1[1]*x2[1], x1[0]*x2[1]+x1[1]*x2[0])
        idx=self.N-idx
        x1=self.v1[idx&self.mask]; x2=self.v2[idx>>self.shift]
        return (x1[0]*x2[0]-x1[1]*x2[1], -(x1[0]*x2[1]+x1[1]*x2[0]))
