#include <stdio.h>
#include <stdlib.h>
#include <math.h>
static double dct8_sel04(double x0,double x1,double x2,double x3,double x4,double x5,double x6,double x7,int sel){
    const double TW3 = 0x1.6a09e667f3bccp-1, HSQ = 0x1.6a09e667f3bcdp-1;
    double c0 = 2.0*x0, c7 = 2.0*x7;
    double c1 = x1+x2; double c3 = x3+x4; double c5 = x5+x6;
    double h0 = c0+c7; double h3 = 2.0*c3; double h1 = c1+c5;
    double a = h0+h3; double e1 = 2.0*h1;
    if (sel==0){ double s = 0.25*(a+e1); return s*HSQ; }
    double s = 0.25*(a-e1); return s*TW3;
}
int main(){
    const double TW3 = 0x1.6a09e667f3bccp-1, HSQ = 0x1.6a09e667f3bcdp-1;
    srand(1); long bad=0, n=0;
    for (long it=0; it<20000000; it++){
        float s[8]; for(int i=0;i<8;i++) s[i]=(float)((rand()%2041)-1024);
        if (it%4==0) { int v=(rand()%2041)-1024; for(int i=0;i<8;i++) s[i]=(float)v; }
        for (int u4=0;u4<2;u4++) for (int v4=0; v4<2; v4++){
            double m = u4?TW3:HSQ;
            double c[8]; for(int i=0;i<8;i++) c[i]=(0.5*(double)s[i])*m;
            double y = dct8_sel04(c[0],c[1],c[2],c[3],c[4],c[5],c[6],c[7], v4?4:0);
            // simplified
            double X[8]; for(int i=0;i<8;i++) X[i]=(double)s[i]*m;
            double A=(X[0]+X[7])+(X[3]+X[4]); double h=(X[1]+X[2])+(X[5]+X[6]);
            double r = v4 ? A-h : A+h;
            double y2 = 0.25*(r*(v4?TW3:HSQ));
            n++; if (y!=y2) { if(bad<5) printf("diff %a %a\n",y,y2); bad++; }
        }
    }
    printf("checked %ld bad %ld\n", n, bad);
    return 0;
}
